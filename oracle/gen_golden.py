"""Generate golden vectors from the UNMODIFIED reference (AniZpZ/AutoSmoothQuant) executed on CPU.

TEST INFRASTRUCTURE.  Run in the build container, where the reference is mounted read-only at
/root/reference:

    python oracle/gen_golden.py            # writes tests/golden/w8a8_golden.npz

The reference's Linear classes import ``autosmoothquant._CUDA`` (a cuBLASLt extension) at module
import time and allocate the int32 output on ``torch.cuda.current_device()``
(autosmoothquant/layers/nn/linear.py:14,101).  To run them on CPU we register a stub ``_CUDA``
module whose ``I8CUGEMM.linear_a8_w8_o32_`` performs the exact integer matmul the cuBLASLt call is
contracted to perform (csrc/int8gemm/bindings.cpp:69-84), and make ``current_device()`` answer
"cpu".  Nothing in the reference is edited or copied; only its outputs are stored.

Stored per case: the inputs (fp32 arrays holding values exactly representable in the case's dtype),
the module buffers, and the reference's output.  16-bit tensors are widened to fp32 (exact).
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

REFERENCE_ROOT = Path("/root/reference")
OUT_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"

DTYPES = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


def import_reference(root=None):
    """Import the reference's layer modules on CPU (see module docstring).  `root` = directory that holds the
    reference's ``autosmoothquant/`` tree: /root/reference here, or the git-ignored copy of its ``layers`` package
    under baseline/_ref on the GPU box (bench.py's reference arm).  With `root` given, only the Linear module is returned."""
    base = Path(root) if root is not None else REFERENCE_ROOT
    if not base.exists():
        raise RuntimeError(f"{base} is not mounted")
    if "autosmoothquant.layers.nn.linear" in sys.modules:
        if root is not None:
            return sys.modules["autosmoothquant.layers.nn.linear"]
        return sys.modules["autosmoothquant.layers.nn.linear"], sys.modules["autosmoothquant.layers.functional.quantization"]
    sys.path.insert(0, str(base))
    stub = types.ModuleType("autosmoothquant._CUDA")

    class I8CUGEMM:  # int8 [M,K] x int8 [N,K]^T -> int32 [M,N], in place (bindings.cpp:69-84)
        def linear_a8_w8_o32_(self, x, w, out):
            out.copy_((x.double() @ w.double().t()).to(torch.int32))

        linear_a8_w8_o32 = linear_a8_w8_o32_

    stub.I8CUGEMM = I8CUGEMM
    sys.modules["autosmoothquant._CUDA"] = stub
    torch.cuda.current_device = lambda: "cpu"  # linear.py:101 allocates `out` on current_device()
    import autosmoothquant.layers.nn.linear as ref_linear  # noqa: E402
    import autosmoothquant.layers.functional.quantization as ref_quant  # noqa: E402
    if root is not None:
        return ref_linear
    return ref_linear, ref_quant


def make_x(gen, shape, dtype, scale=1.0, outlier=True, zero_row=True):
    x = torch.randn(*shape, generator=gen) * scale
    if outlier and shape[-1] >= 8:
        x[..., 3] *= 25.0
    x2 = x.reshape(-1, shape[-1])
    if zero_row and x2.shape[0] > 1:
        x2[1].zero_()
    if x2.shape[0] > 2:
        x2[2, 0] = 1e4  # saturates in per-tensor modes
        x2[2, 1] = -1e4
    return x.to(dtype)


def f32(t: torch.Tensor) -> np.ndarray:
    return t.detach().to(torch.float32).cpu().numpy()


def main() -> None:
    ref_linear, ref_quant = import_reference()
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    arrays = {}
    meta = {"generator": "oracle/gen_golden.py", "torch": torch.__version__, "cases": []}
    gen = torch.Generator().manual_seed(20240607)

    def add(case_id, kind, **kw):
        entry = {"id": case_id, "kind": kind}
        for k, v in kw.items():
            if isinstance(v, (np.ndarray, torch.Tensor)):
                a = v.numpy() if isinstance(v, torch.Tensor) else v
                arrays[f"{case_id}.{k}"] = a
            else:
                entry[k] = v
        meta["cases"].append(entry)

    cid = 0
    shapes = [((5, 64), 48), ((2, 7, 96), 40), ((1, 32), 16), ((130, 160), 24)]
    # ---- INT8 Linear / LinearWithQuantScale (linear.py:35-129, 248-329)
    for dname, dtype in DTYPES.items():
        for (xs, N) in shapes:
            K = xs[-1]
            for cls_name in ("W8A8BFP32OFP32Linear", "W8A8BFP32OFP32LinearWithQuantScale"):
                for act_quant in ("per-tensor", "per-token"):
                    for use_bias in (False, True):
                        lin = torch.nn.Linear(K, N, bias=use_bias)
                        with torch.no_grad():
                            lin.weight.copy_(torch.randn(N, K, generator=gen) * 0.05)
                            if use_bias:
                                lin.bias.copy_(torch.randn(N, generator=gen))
                        input_scale = 0.0371
                        cls = getattr(ref_linear, cls_name)
                        mod = cls.from_float(lin, input_scale, act_quant=act_quant)
                        # per-tensor Linear expects pre-scaled ("int8 units") input, the others raw activations
                        pre_scaled = cls_name == "W8A8BFP32OFP32Linear" and act_quant == "per-tensor"
                        x = make_x(gen, xs, dtype, scale=30.0 if pre_scaled else 1.0)
                        y = mod(x)
                        assert y.dtype == dtype and y.shape == (*xs[:-1], N)
                        kw = dict(
                            x=f32(x), weight=mod.weight.numpy().copy(), y=f32(y),
                            dequant_scale=float(mod.dequant_scale), dtype=dname, cls=cls_name,
                            act_quant=act_quant,
                        )
                        if use_bias:
                            kw["bias"] = mod.bias.detach().numpy().copy()
                        if hasattr(mod, "quant_scale"):
                            kw["quant_scale"] = float(mod.quant_scale)
                        add(f"c{cid}", "int8_linear", **kw)
                        cid += 1
    # ---- QKV variant (linear.py:132-245)
    for dname, dtype in DTYPES.items():
        for act_quant in ("per-tensor", "per-token"):
            for use_bias in (False, True):
                K, qkv = 64, [32, 16, 16]
                lin = torch.nn.Linear(K, sum(qkv), bias=use_bias)
                with torch.no_grad():
                    lin.weight.copy_(torch.randn(sum(qkv), K, generator=gen) * torch.tensor([0.03] * 32 + [0.06] * 16 + [0.1] * 16).view(-1, 1))
                    if use_bias:
                        lin.bias.copy_(torch.randn(sum(qkv), generator=gen))
                mod = ref_linear.W8A8BFP32OFP32QKVLinear.from_float(lin, 0.0412, qkv, act_quant=act_quant)
                x = make_x(gen, (3, 5, K), dtype, scale=30.0 if act_quant == "per-tensor" else 1.0)
                y = mod(x)
                kw = dict(x=f32(x), weight=mod.weight.numpy().copy(), y=f32(y), dtype=dname, act_quant=act_quant,
                          qkv_size=qkv, q_scale=float(mod.q_dequant_scale), k_scale=float(mod.k_dequant_scale),
                          v_scale=float(mod.v_dequant_scale))
                if use_bias:
                    kw["bias"] = mod.bias.detach().numpy().copy()
                add(f"c{cid}", "int8_qkv", **kw)
                cid += 1
    # ---- offline weight quantiser (quantization.py:9-18)
    for dname, dtype in DTYPES.items():
        w = (torch.randn(24, 40, generator=gen) * 0.07).to(dtype)
        w_in = f32(w)
        q, s = ref_quant.quantize_per_tensor_absmax(w.clone())
        add(f"c{cid}", "weight_quant", w=w_in, q=q.numpy().copy(), scale=float(s), dtype=dname)
        cid += 1
    # ---- FP8 activation quantisers (quantization.py:144-211)
    for dname, dtype in DTYPES.items():
        x = make_x(gen, (9, 64), dtype)
        x[2, 0] = 300.0
        x[2, 1] = -300.0
        xr = x.clone()
        xr[1] = torch.randn(64, generator=gen).to(dtype) * 1e-3  # per-token asserts/NaNs on zero rows: keep one tiny row
        q, s = ref_quant.per_token_quantize_fp8(xr)
        add(f"c{cid}", "fp8_per_token", x=f32(xr), q=q.view(torch.uint8).numpy().copy(), scale=f32(s).reshape(-1), dtype=dname)
        cid += 1
        q = ref_quant.static_per_tensor_quantize_fp8(x, 0.55)
        add(f"c{cid}", "fp8_static", x=f32(x), q=q.view(torch.uint8).numpy().copy(), in_scale=0.55, dtype=dname)
        cid += 1
        q, s = ref_quant.per_tensor_quantize_fp8(x)
        add(f"c{cid}", "fp8_per_tensor", x=f32(x), q=q.view(torch.uint8).numpy().copy(), scale=float(s), dtype=dname)
        cid += 1
    # ---- FP8 linears on fp32 activations (linear.py:371-452, 502-581); the reference only runs fp8 in fp32
    for act in ("per-token", "static"):
        for use_bias in (False, True):
            K, N = 64, 40
            wf = torch.randn(N, K, generator=gen) * 0.05
            wq, wscale = ref_quant.per_tensor_quantize_fp8(wf)
            bias = torch.randn(N, generator=gen) if use_bias else None
            x = make_x(gen, (11, K), torch.float32, zero_row=False)
            if act == "per-token":
                mod = ref_linear.FP8LinearDynamic(K, N, "per-token", use_bias)
            else:
                mod = ref_linear.FP8LinearStatic(K, N, use_bias)
                mod.input_scale = torch.tensor(float(x.abs().max() / 448.0))
                mod.output_scale = torch.tensor(0.0)  # falsy -> no output fake-quant (linear.py:562)
            mod.weight = wq
            mod.weight_scale = wscale.to(torch.float32)
            if use_bias:
                mod.bias = bias
            y = mod(x)
            kw = dict(x=f32(x), w=wq.view(torch.uint8).numpy().copy(), y=f32(y), w_scale=float(wscale), act=act,
                      dtype="f32")
            if act == "static":
                kw["in_scale"] = float(mod.input_scale)
            if use_bias:
                kw["bias"] = bias.numpy().copy()
            add(f"c{cid}", "fp8_linear", **kw)
            cid += 1

    # ---- FP8LinearDynamic per-tensor dynamic branch (linear.py:417-418) and FP8LinearStatic with output
    #      fake-quantisation (linear.py:562-564)
    for dname in ("f32", "bf16"):
        dtype = DTYPES[dname]
        K, N = 64, 40
        wf = torch.randn(N, K, generator=gen) * 0.05
        wq, wscale = ref_quant.per_tensor_quantize_fp8(wf)
        bias = torch.randn(N, generator=gen)
        x = make_x(gen, (11, K), dtype, zero_row=False)
        mod = ref_linear.FP8LinearDynamic(K, N, "per-tensor", True)
        mod.weight = wq
        mod.weight_scale = wscale.to(torch.float32)
        mod.bias = bias.to(dtype)
        y = mod(x)
        add(f"c{cid}", "fp8_linear", x=f32(x), w=wq.view(torch.uint8).numpy().copy(), y=f32(y), w_scale=float(wscale),
            act="per-tensor", dtype=dname, bias=f32(bias.to(dtype)))
        cid += 1
    for out_scale in (0.031, 0.2):
        K, N = 64, 40
        wf = torch.randn(N, K, generator=gen) * 0.05
        wq, wscale = ref_quant.per_tensor_quantize_fp8(wf)
        x = make_x(gen, (11, K), torch.float32, zero_row=False)
        mod = ref_linear.FP8LinearStatic(K, N, False)
        mod.weight = wq
        mod.weight_scale = wscale.to(torch.float32)
        mod.input_scale = torch.tensor(float(x.abs().max() / 448.0))
        mod.output_scale = torch.tensor(out_scale)
        y = mod(x)
        add(f"c{cid}", "fp8_linear", x=f32(x), w=wq.view(torch.uint8).numpy().copy(), y=f32(y), w_scale=float(wscale),
            act="static", in_scale=float(mod.input_scale), out_scale=out_scale, dtype="f32")
        cid += 1

    np.savez_compressed(OUT_DIR / "w8a8_golden.npz", **arrays)
    (OUT_DIR / "w8a8_golden.json").write_text(json.dumps(meta, indent=1))
    size = (OUT_DIR / "w8a8_golden.npz").stat().st_size
    print(f"wrote {cid} cases, {len(arrays)} arrays, {size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
