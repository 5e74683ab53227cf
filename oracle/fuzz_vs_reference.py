"""Fuzz the numpy oracle against the LIVE, unmodified reference (CPU), beyond the committed golden vectors.

TEST INFRASTRUCTURE.  Needs /root/reference (the build container); run as its own process because importing the
reference registers its modules under the ``autosmoothquant.*`` names and redirects ``torch.cuda.current_device``:

    python oracle/fuzz_vs_reference.py --cases 300 --seed 0        # prints one JSON summary line

Every case draws a shape (M 1..70 incl. 1, K a multiple of 16 up to 272, N 8..96), an activation dtype, a module class
(W8A8BFP32OFP32Linear / ...WithQuantScale / ...QKVLinear, built by the reference's own ``from_float`` from a random
nn.Linear), a granularity and a bias flag, runs the reference's ``forward`` with the exact-integer ``_CUDA`` stub, and
requires the oracle (``w8a8_oracle.w8a8_linear`` / ``w8a8_qkv_linear``, division mode "exact" = torch on CPU) to
reproduce the output BIT FOR BIT; the weight quantiser and the three FP8 activation quantisers are compared the same way.
Inputs carry outlier channels, saturating values, an all-zero row and ties at .5 (per-tensor modes).
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import w8a8_oracle as O  # noqa: E402
from oracle.gen_golden import DTYPES, f32, import_reference  # noqa: E402


def make_input(g, M, K, dtype, scale):
    x = torch.randn(M, K, generator=g) * scale
    if K >= 8:
        x[:, int(torch.randint(0, K, (1,), generator=g))] *= 25.0  # an outlier channel
    if M > 1:
        x[1].zero_()  # all-zero token
    if M > 2:
        x[2, 0], x[2, 1] = 1e4, -1e4  # saturates per-tensor codes
    if M > 3:
        x[3, : min(K, 8)] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.5, -3.5])[: min(K, 8)] * scale  # ties
    return x.to(dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    ref_linear, ref_quant = import_reference()
    g = torch.Generator().manual_seed(args.seed)
    rnd = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))  # noqa: E731
    names = list(DTYPES)
    mismatches, done = [], {"int8_linear": 0, "int8_qkv": 0, "weight_quant": 0, "fp8_quant": 0}
    for case in range(args.cases):
        dname = names[rnd(0, 2)]
        dtype = DTYPES[dname]
        M, K, N = rnd(1, 70), 16 * rnd(1, 17), rnd(8, 96)
        kind = ("linear", "quantscale", "qkv", "fp8q")[rnd(0, 3)]
        act = ("per-tensor", "per-token")[rnd(0, 1)]
        use_bias = bool(rnd(0, 1))
        tag = f"#{case} {kind} {act} {dname} M{M} K{K} N{N} bias{int(use_bias)}"
        lin = torch.nn.Linear(K, N, bias=use_bias)
        with torch.no_grad():
            lin.weight.copy_(torch.randn(N, K, generator=g) * 0.05)
            if use_bias:
                lin.bias.copy_(torch.randn(N, generator=g))
        # weight quantiser on a dtype-rounded copy (the reference quantises the module's weight in place)
        w_t = lin.weight.detach().clone().to(dtype)
        q_ref, s_ref = ref_quant.quantize_per_tensor_absmax(w_t.clone())
        q_or, s_or = O.quantize_per_tensor_absmax(f32(w_t), dname)
        done["weight_quant"] += 1
        if float(s_ref) != float(s_or) or not np.array_equal(q_ref.numpy(), q_or):
            mismatches.append(tag + " weight_quant")
        if kind == "fp8q":
            x = make_input(g, M, K, dtype, 3.0)
            for mode, fn, extra in (("per-token", ref_quant.per_token_quantize_fp8, ()),
                                    ("per-tensor", ref_quant.per_tensor_quantize_fp8, ()),
                                    ("scale", ref_quant.static_per_tensor_quantize_fp8, (0.37,))):
                out = fn(x, *extra)
                q_t = out[0] if isinstance(out, tuple) else out
                q_o, s_o = O.quantize_act_fp8(f32(x), dname, mode, extra[0] if extra else 1.0, div_mode="exact")
                want = q_t.view(torch.uint8).numpy()
                nan = (want & 0x7F) == 0x7F  # 0/0 rows: NaN payload sign is library-specific
                ok = np.array_equal(((q_o & 0x7F) == 0x7F), nan) and np.array_equal(q_o[~nan], want[~nan])
                if ok and mode == "per-token":
                    ok = np.array_equal(np.asarray(s_o, np.float32).reshape(-1), out[1].float().numpy().reshape(-1))
                if ok and mode == "per-tensor":
                    ok = float(np.asarray(s_o).reshape(())) == float(out[1])
                done["fp8_quant"] += 1
                if not ok:
                    mismatches.append(f"{tag} fp8 {mode}")
            continue
        scale = 30.0 if (kind != "quantscale" and act == "per-tensor") else 1.0
        x = make_input(g, M, K, dtype, scale)
        if kind == "qkv":
            n1 = max(8, (N // 3) // 8 * 8)
            sizes = [N - 2 * n1, n1, n1] if N - 2 * n1 > 0 else [N, 0, 0]
            if sizes[1] == 0:
                continue
            mod = ref_linear.W8A8BFP32OFP32QKVLinear.from_float(lin, 0.04, sizes, act_quant=act)
            y = mod(x)
            got = O.w8a8_qkv_linear(f32(x), dname, mod.weight.numpy(), sizes, float(mod.q_dequant_scale), float(mod.k_dequant_scale),
                                    float(mod.v_dequant_scale), act_quant=act, bias=mod.bias.detach().numpy() if use_bias else None, div_mode="exact")
            done["int8_qkv"] += 1
        else:
            cls = ref_linear.W8A8BFP32OFP32Linear if kind == "linear" else ref_linear.W8A8BFP32OFP32LinearWithQuantScale
            mod = cls.from_float(lin, 0.04, act_quant=act)
            y = mod(x)
            qs = float(mod.quant_scale) if (kind == "quantscale" and act == "per-tensor") else None
            got = O.w8a8_linear(f32(x), dname, mod.weight.numpy(), float(mod.dequant_scale), act_quant=act,
                                bias=mod.bias.detach().numpy() if use_bias else None, quant_scale=qs, div_mode="exact")
            done["int8_linear"] += 1
        want = f32(y)
        if y.dtype != dtype or want.shape != got.shape or not np.array_equal(want, got, equal_nan=True):
            mismatches.append(tag)
    print(json.dumps({"cases": args.cases, "seed": args.seed, "compared": done, "mismatches": mismatches[:20],
                      "n_mismatches": len(mismatches), "torch": torch.__version__}))
    return 1 if mismatches else 0


if __name__ == "__main__":
    sys.exit(main())
