#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02d
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/debug_nvls.py check > gpurun_out/r02d/debug_nvls_w2.log 2>&1
echo "rc=$?" >> gpurun_out/r02d/debug_nvls_w2.log
for r in 8 16 24 32; do
ASQ_NVLS_REDUCERS=$r timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 scripts/perf_allreduce.py nvls-only > gpurun_out/r02d/perf_allreduce_w2_red$r.log 2>&1
done
timeout 600 python -m pytest tests/test_fused_allreduce.py -x -q -m gpu > gpurun_out/r02d/tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02d/tests.log
grep -h "world" gpurun_out/r02d/perf_allreduce_w2*.log | cut -c1-400
grep -c "!= fp32-sum 0/" gpurun_out/r02d/debug_nvls_w2.log
tail -5 gpurun_out/r02d/tests.log
