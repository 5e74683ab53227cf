#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02i
ASQ_NVLS_REDUCERS=32 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/debug_nvls.py timeline > gpurun_out/r02i/timeline_w2.log 2>&1
grep -A16 "rank 0 2048" gpurun_out/r02i/timeline_w2.log | grep -E "rank 0|signalling|last tile stored|reducer: |first slab landed|handshake"
grep -A16 "rank 0 4096" gpurun_out/r02i/timeline_w2.log | grep -E "rank 0|signalling|last tile stored|reducer: |first slab landed|handshake"
