#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02j
ASQ_NVLS_REDUCERS=32 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/debug_nvls.py timeline check > gpurun_out/r02j/timeline_w2.log 2>&1
grep -A16 "rank 0 2048" gpurun_out/r02j/timeline_w2.log | grep -E "^rank 0 2048x4096x4096: k|signalling|last tile stored|reducer: |first slab landed|handshake"
grep  "rank 0 .*#0" gpurun_out/r02j/timeline_w2.log | cut -c1-120
for r in 16 32; do
ASQ_NVLS_REDUCERS=$r timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 scripts/perf_allreduce.py nvls-only > gpurun_out/r02j/perf_allreduce_w2_red$r.log 2>&1
done
grep -h "world" gpurun_out/r02j/perf_allreduce_w2*.log | cut -c1-400
timeout 600 python -m pytest tests/test_fused_allreduce.py tests/test_tp_nccl.py -x -q -m gpu > gpurun_out/r02j/tests.log 2>&1
tail -3 gpurun_out/r02j/tests.log
