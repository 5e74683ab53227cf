#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02g
for dbgv in 0x0 0x100 0x200 0x400 0x600 0x700; do
ASQ_NVLS_DEBUG=$dbgv timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/debug_nvls.py timeline > gpurun_out/r02g/timeline_w2_$dbgv.log 2>&1
echo "== ASQ_NVLS_DEBUG=$dbgv"
grep -A12 "rank 0 2048" gpurun_out/r02g/timeline_w2_$dbgv.log | grep -E "rank 0|signalling|last tile stored|reducer: done|first slab landed"
done
