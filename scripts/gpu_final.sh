#!/bin/bash
# Last bounded GPU pass of round 2 (1 GPU):  gpurun --timeout 100 -- 'bash scripts/gpu_final.sh r02p 82'
# full GPU test suite on the rebuilt library, then A/B bench lines (no secondary measurements, no CPU baseline):
#   o_proj quantised in-kernel (default) vs stand-alone quantise kernel + int8-in GEMM (ASQ_OPROJ_SPLIT=1)
#   FP8 per-token Llama-2-7B: fused q|k|v / gate|up launches (default) vs one launch per projection (ASQ_FP8_FUSE=0)
TAG=${1:-final}
LIMIT=${2:-82}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
START=$(date +%s)
left() { echo $(( LIMIT - ($(date +%s) - START) )); }
run() {
  local name=$1 cap=$2; shift 2
  local l; l=$(left)
  if [ "$l" -lt 8 ]; then echo "$name skipped (${l}s left)"; return; fi
  local t=$(( cap < l ? cap : l ))
  local t0; t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  echo "$name rc=$? in $(( $(date +%s) - t0 ))s (cap ${t}s)"
}
B="python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline"
FP8="--quant type=fp8,qkv=per-token,out=per-token,fc1=per-token,fc2=per-token"
run pytest_gpu 60 python -m pytest tests -m gpu -x -q
tail -3 "$OUT/pytest_gpu.log"
run bench_default 30 env ASQ_OPROJ_SPLIT=0 $B
run bench_oproj_split 30 env ASQ_OPROJ_SPLIT=1 $B
run bench_fp8_fused 30 $B $FP8
run bench_fp8_unfused 30 env ASQ_FP8_FUSE=0 $B $FP8
run bench_oproj_split_pdl 30 env ASQ_OPROJ_SPLIT=1 ASQ_PDL=1 $B
echo "total $(( $(date +%s) - START ))s"
python - "$OUT" <<'EOF'
import json, sys, glob, os
for f in sorted(glob.glob(os.path.join(sys.argv[1], "bench_*.log"))):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        r = d["roofline"]
        print(os.path.basename(f), "ms/step %.3f" % d["ms_per_step"], "tok/s %.0f" % d["value"], "linears TOPS %.0f" % (r["achieved"] or 0),
              "launches/step", d["gpu_launches"] // d["steps"],
              " | ".join("%s %dx%dx%d %.1fus" % (b["entry"][:18], b["M"], b["N"], b["K"], b["avg_us"]) for b in r["by_launch_shape"]))
    except Exception as e:
        print(os.path.basename(f), "unreadable:", repr(e)[:100])
EOF
