#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02h
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/debug_nvls.py timeline check > gpurun_out/r02h/timeline_w2.log 2>&1
grep -A12 "rank 0 2048" gpurun_out/r02h/timeline_w2.log | grep -E "rank 0|signalling|last tile stored|reducer: |first slab landed"
grep -c "!= fp32-sum" gpurun_out/r02h/timeline_w2.log
for r in 16 32 48; do
ASQ_NVLS_REDUCERS=$r timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 scripts/perf_allreduce.py nvls-only > gpurun_out/r02h/perf_allreduce_w2_red$r.log 2>&1
done
grep -h "world" gpurun_out/r02h/perf_allreduce_w2*.log | cut -c1-400
timeout 600 python -m pytest tests/test_fused_allreduce.py -x -q -m gpu > gpurun_out/r02h/tests.log 2>&1
tail -3 gpurun_out/r02h/tests.log
