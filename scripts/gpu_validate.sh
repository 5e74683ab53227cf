#!/bin/bash
# One bounded validation pass on a 1-GPU box:  gpurun --timeout 285 -- 'bash scripts/gpu_validate.sh r02n 262'
# Every step runs under `timeout` capped by what is left of the overall limit, most important first; logs land in
# gpurun_out/<tag>/ (merged back by gpurun), a one-line verdict per step goes to stdout.
TAG=${1:-validate}
LIMIT=${2:-262}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
START=$(date +%s)
left() { echo $(( LIMIT - ($(date +%s) - START) )); }
run() {
  local name=$1 cap=$2; shift 2
  local l; l=$(left)
  if [ "$l" -lt 12 ]; then echo "$name skipped (${l}s left)"; return; fi
  local t=$(( cap < l ? cap : l ))
  local t0; t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  echo "$name rc=$? in $(( $(date +%s) - t0 ))s (cap ${t}s)"
}
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > "$OUT/gpu.txt" 2>&1
run pytest_gpu 170 python -m pytest tests -m gpu -x -q
tail -4 "$OUT/pytest_gpu.log"
run bench_n1 130 python bench.py --steps 10 --warmup 3
run smoke 40 python -c "import __graft_entry__ as g; g.smoke()"
tail -1 "$OUT/smoke.log"
run bench_cfg4_n1 120 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary
run ceiling 50 python scripts/int8_ceiling.py --write
cp profiles/int8_ceiling.json "$OUT/" 2>/dev/null
echo "total $(( $(date +%s) - START ))s"
head -c 600 "$OUT/bench_n1.log"; echo
head -c 400 "$OUT/bench_cfg4_n1.log"; echo
