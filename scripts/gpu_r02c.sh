#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02c
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/debug_nvls.py check > gpurun_out/r02c/debug_nvls_w2.log 2>&1
echo "rc=$?" >> gpurun_out/r02c/debug_nvls_w2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 scripts/perf_allreduce.py > gpurun_out/r02c/perf_allreduce_w2.log 2>&1
for r in 4 8 32; do
ASQ_NVLS_REDUCERS=$r timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 scripts/perf_allreduce.py nvls-only > gpurun_out/r02c/perf_allreduce_w2_red$r.log 2>&1
done
timeout 600 python -m pytest tests/test_fused_allreduce.py tests/test_tp_nccl.py tests/test_moe_grouped.py -x -q -m gpu > gpurun_out/r02c/tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02c/tests.log
grep -h "world" gpurun_out/r02c/perf_allreduce_w2*.log | cut -c1-400
tail -5 gpurun_out/r02c/tests.log
