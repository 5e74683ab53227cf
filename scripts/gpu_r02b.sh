#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02b
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/debug_nvls.py > gpurun_out/r02b/debug_nvls_w2.log 2>&1
echo "rc=$?" >> gpurun_out/r02b/debug_nvls_w2.log
tail -40 gpurun_out/r02b/debug_nvls_w2.log
