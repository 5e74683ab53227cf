#!/usr/bin/env python
"""tcgen05 issue-rate ceiling of the 8-bit tensor pipe: runs csrc/asq_ceiling.cu on cuda:0.

    python scripts/int8_ceiling.py [--write] [--groups 4096] [--reps 20]

Every variant (kind::i8 / kind::f8f6f4 x cta_group 1 / 2) runs in its own subprocess with a timeout, so a variant
that fails cannot take the others (or the caller: bench.py) with it.  Prints one JSON object; --write also stores it as
profiles/int8_ceiling.json.  The numbers are the denominators of `roofline` in bench.py: MMAs of the kernel's own
shape issued back to back from resident shared-memory operands — no TMA, no epilogue, no HBM traffic.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "autosmoothquant_b200" / "libasq_ceiling.so"
VARIANTS = [("i8", 2), ("i8", 1), ("fp8", 2), ("fp8", 1)]


def run_one(kind: str, cta_group: int, groups: int, reps: int) -> dict:
    lib = ctypes.CDLL(str(LIB))
    lib.asq_mma_ceiling.restype = ctypes.c_int
    lib.asq_mma_ceiling.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    res = (ctypes.c_double * 6)()
    rc = lib.asq_mma_ceiling(1 if kind == "fp8" else 0, cta_group, groups, reps, res)
    out = {"kind": kind, "cta_group": cta_group, "rc": rc}
    if rc == 0:
        out.update({"tops": res[0], "cycles_per_mma": res[1], "sm_mhz_effective": res[2], "ms_per_launch": res[3],
                    "ctas": int(res[4]), "ops_per_mma": res[5], "launches": reps,
                    # per SM: MACs retired per clock = (ops / 2) per MMA / cycles per MMA / SMs per MMA
                    "macs_per_clk_per_sm": (res[5] / 2.0) / res[1] / cta_group if res[1] else None})
    return out


def measure(groups: int = 4096, reps: int = 20, timeout: float = 30.0) -> dict:
    if not LIB.exists():
        return {"unavailable": f"{LIB.name} not built (autosmoothquant_b200.build.build_ceiling)"}
    results = []
    for kind, cg in VARIANTS:
        cmd = [sys.executable, str(Path(__file__).resolve()), "--one", kind, str(cg), "--groups", str(groups), "--reps", str(reps)]
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
            line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            results.append(json.loads(line[-1]) if line else {"kind": kind, "cta_group": cg, "rc": p.returncode, "stderr": p.stderr[-300:]})
        except subprocess.TimeoutExpired:
            results.append({"kind": kind, "cta_group": cg, "rc": "timeout"})
        except Exception as e:  # noqa: BLE001
            results.append({"kind": kind, "cta_group": cg, "rc": repr(e)[:200]})
    ok = [r for r in results if r.get("rc") == 0]
    best = {}
    for kind in ("i8", "fp8"):
        cand = [r["tops"] for r in ok if r["kind"] == kind]
        if cand:
            best[kind] = max(cand)
    return {"what": "tcgen05.mma issue rate from resident smem operands, 128x256x32 (cta_group::1) / 256x256x32 (cta_group::2) "
                    "MMAs back to back on every SM, two groups of 64 in flight, CUDA-event timed",
            "unit": "TOP/s (i8) / TFLOP/s (fp8)", "variants": results, "best": best, "measured_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", nargs=2, metavar=("KIND", "CTA_GROUP"))
    ap.add_argument("--groups", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--write", action="store_true")
    args = ap.parse_args()
    if args.one:
        print(json.dumps(run_one(args.one[0], int(args.one[1]), args.groups, args.reps)), flush=True)
        return
    out = measure(args.groups, args.reps)
    print(json.dumps(out, indent=1))
    if args.write and out.get("best"):
        (ROOT / "profiles" / "int8_ceiling.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
