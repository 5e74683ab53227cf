#!/bin/bash
# round-2 first multi-GPU check (2 GPUs): new tests, all-reduce perf, TP bench with parity gate
set -u
mkdir -p gpurun_out/r02a
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_fused_allreduce.py tests/test_tp_nccl.py tests/test_moe_grouped.py -x -q -m gpu > gpurun_out/r02a/tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02a/tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/perf_allreduce.py > gpurun_out/r02a/perf_allreduce_w2.log 2>&1
echo "rc=$?" >> gpurun_out/r02a/perf_allreduce_w2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02a/bench_tp2_auto.json 2> gpurun_out/r02a/bench_tp2_auto.err
echo "rc=$?" >> gpurun_out/r02a/bench_tp2_auto.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --tp-reduce nccl --no-secondary --no-parity > gpurun_out/r02a/bench_tp2_nccl.json 2> gpurun_out/r02a/bench_tp2_nccl.err
echo "rc=$?" >> gpurun_out/r02a/bench_tp2_nccl.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --tp-reduce fused --no-secondary --no-parity > gpurun_out/r02a/bench_tp2_fused.json 2> gpurun_out/r02a/bench_tp2_fused.err
echo "rc=$?" >> gpurun_out/r02a/bench_tp2_fused.err
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a/bench_n1.json 2> gpurun_out/r02a/bench_n1.err
echo "rc=$?" >> gpurun_out/r02a/bench_n1.err
tail -3 gpurun_out/r02a/tests.log
