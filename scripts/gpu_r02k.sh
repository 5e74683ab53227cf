#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02k
run() { name=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 "$@" > gpurun_out/r02k/$name.json 2> gpurun_out/r02k/$name.err; echo "$name rc=$?"; tail -c 600 gpurun_out/r02k/$name.err | grep -v "^\*\|OMP_NUM" | tail -5; }
run bench_tp2_cfg4_l2 --config 4 --layers 2 --steps 5 --warmup 3
run bench_tp2_cfg5_l2 --config 5 --layers 2 --steps 5 --warmup 3
run bench_tp2_cfg5_l2_nccl --config 5 --layers 2 --steps 5 --warmup 3 --tp-reduce nccl
run bench_tp2_nvls --steps 10 --warmup 3 --tp-reduce nvls --no-secondary
run bench_tp2_default --steps 10 --warmup 3
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02k/tests_all.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r02k/tests_all.log
