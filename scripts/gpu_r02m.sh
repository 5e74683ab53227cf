#!/bin/bash
# 8-GPU run: fused all-reduce tests + perf, NVLS probe, TP8 bench (default / nccl), configs 4 and 5
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r02m; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 400 python -m pytest tests/test_fused_allreduce.py -x -q -m gpu -k "8" > $O/tests_w8.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests_w8.log
PROBE_MB=128 timeout 200 $TR --master-port 29541 scripts/debug_nvls.py probe > $O/probe_w8.log 2>&1; grep -E "^probe|ncclAll" $O/probe_w8.log | cut -c1-200
timeout 200 $TR --master-port 29542 scripts/debug_nvls.py timeline > $O/timeline_w8.log 2>&1; grep -A16 "^rank 0 " $O/timeline_w8.log | grep -E "^rank 0|signalling|last tile|reducer: |handshake" | cut -c1-200
timeout 300 $TR --master-port 29543 scripts/perf_allreduce.py > $O/perf_allreduce_w8.log 2>&1; grep -h "^world" $O/perf_allreduce_w8.log | cut -c1-420
for r in 8 16 64; do ASQ_NVLS_REDUCERS=$r timeout 200 $TR --master-port 29544 scripts/perf_allreduce.py nvls-only > $O/perf_allreduce_w8_red$r.log 2>&1; grep -h "^world" $O/perf_allreduce_w8_red$r.log | cut -c1-200; done
run() { name=$1; shift; timeout 700 $TR --master-port 29545 bench.py --gpus 8 "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; grep -v "^\*\|OMP_NUM\|^$" $O/$name.err | tail -4; cut -c1-400 $O/$name.json; echo; }
run bench_tp8_default --steps 10 --warmup 3
run bench_tp8_nccl --steps 10 --warmup 3 --tp-reduce nccl --no-secondary --no-parity
run bench_tp8_cfg5 --config 5 --steps 5 --warmup 3
run bench_tp8_cfg4 --config 4 --steps 5 --warmup 3
