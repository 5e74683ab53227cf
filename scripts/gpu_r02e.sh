#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02e
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/debug_nvls.py timeline > gpurun_out/r02e/timeline_w2.log 2>&1
echo "rc=$?" >> gpurun_out/r02e/timeline_w2.log
grep -v "^$" gpurun_out/r02e/timeline_w2.log | tail -60
