"""Golden state dicts written by the UNMODIFIED reference modules (TEST INFRASTRUCTURE).

    python oracle/gen_golden_state.py        # writes tests/golden/state_golden.{npz,json}

What the reference puts on disk is `state_dict()` of its quantized modules (`save_pretrained`,
examples/smoothquant_model.py:96-99): int8 / e4m3 `weight`, fp32 `bias`, 0-dim fp32 scale buffers.  For every module
class this script builds the reference module the way the reference does (from_float converters; for FP8 also the
constructor + buffer assignment its model classes perform when they load a checkpoint, models/llama.py:83-90), stores
its complete state dict, one input and the reference's output on CPU.  A decoder-layer-shaped dict uses the submodule
names and class choices of Int8LlamaAttention / Int8LlamaMLP (models/llama.py:74-106, 185-214) for one quant_config.
tests/test_state_dict_compat.py loads these into the B200 modules with load_state_dict(strict=True).
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.gen_golden import import_reference, make_x  # noqa: E402

OUT_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"


def to_np(t: torch.Tensor):
    """(array, dtype name): float8 as its bytes, 16-bit floats widened (exact)."""
    t = t.detach().cpu()
    if t.dtype == torch.float8_e4m3fn:
        return t.view(torch.uint8).numpy().copy(), "float8_e4m3fn"
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.float().numpy().copy(), str(t.dtype).replace("torch.", "")
    return t.numpy().copy(), str(t.dtype).replace("torch.", "")


def main() -> None:
    L, Q = import_reference()
    gen = torch.Generator().manual_seed(20241017)
    arrays, cases = {}, []

    def lin(i, o, bias, std=0.05):
        m = torch.nn.Linear(i, o, bias=bias)
        with torch.no_grad():
            m.weight.copy_(torch.randn(o, i, generator=gen) * std)
            if bias:
                m.bias.copy_(torch.randn(o, generator=gen))
        return m

    def add(cid, mod, meta, x=None, y=None):
        sd = mod.state_dict()
        entry = dict(meta, id=cid, keys=[], dtypes={}, shapes={})
        for k, v in sd.items():
            a, dt = to_np(v)
            arrays[f"{cid}.sd.{k}"] = a
            entry["keys"].append(k)
            entry["dtypes"][k] = dt
            entry["shapes"][k] = list(v.shape)
        if x is not None:
            arrays[f"{cid}.x"], entry["x_dtype"] = to_np(x)
            arrays[f"{cid}.y"], _ = to_np(y)
        cases.append(entry)

    n = 0
    K, N = 64, 48
    for cls_name in ("W8A8BFP32OFP32Linear", "W8A8BFP32OFP32LinearWithQuantScale"):
        for act in ("per-tensor", "per-token"):
            for bias in (False, True):
                mod = getattr(L, cls_name).from_float(lin(K, N, bias), 0.0371, act_quant=act)
                pre = cls_name == "W8A8BFP32OFP32Linear" and act == "per-tensor"
                x = make_x(gen, (2, 5, K), torch.bfloat16, scale=30.0 if pre else 1.0)
                add(f"s{n}", mod, dict(cls=cls_name, ctor=dict(in_features=K, out_features=N, use_bias=bias, act_quant=act)), x, mod(x))
                n += 1
    for act in ("per-tensor", "per-token"):
        qkv = [32, 16, 16]
        mod = L.W8A8BFP32OFP32QKVLinear.from_float(lin(K, sum(qkv), True), 0.0412, qkv, act_quant=act)
        x = make_x(gen, (3, K), torch.bfloat16, scale=30.0 if act == "per-tensor" else 1.0)
        add(f"s{n}", mod, dict(cls="W8A8BFP32OFP32QKVLinear",
                               ctor=dict(qkv_size=qkv, in_features=K, out_features=sum(qkv), use_bias=True, act_quant=act)), x, mod(x))
        n += 1
    # FP8 dynamic: the converter's own product (per-tensor branch, stored-but-unused bias: linear.py:444-451) ...
    for bias in (False, True):
        mod = L.FP8LinearDynamic.from_float(lin(K, N, bias), 1.0)
        x = make_x(gen, (7, K), torch.float32, zero_row=False)
        add(f"s{n}", mod, dict(cls="FP8LinearDynamic", how="from_float", ctor=dict(in_features=K, out_features=N, act_quant=mod.act_quant,
                                                                                 use_bias=mod.use_bias)), x, mod(x))
        n += 1
    # ... and the module a model constructor builds and fills from the checkpoint (per-token, bias used)
    for bias in (False, True):
        src = L.FP8LinearDynamic.from_float(lin(K, N, bias), 1.0)
        mod = L.FP8LinearDynamic(K, N, "per-token", bias)
        mod.weight, mod.weight_scale = src.weight, src.weight_scale
        if bias:
            mod.bias = src.bias.data.to(torch.float32)
        x = make_x(gen, (7, K), torch.float32, zero_row=False)
        add(f"s{n}", mod, dict(cls="FP8LinearDynamic", how="ctor", ctor=dict(in_features=K, out_features=N, act_quant="per-token", use_bias=bias)),
            x, mod(x))
        n += 1
    # FP8 static: calibrate with the reference's observer, convert (linear.py:455-499, 568-581)
    for bias in (False, True):
        fl = lin(K, N, bias)
        wq, ws = Q.per_tensor_quantize_fp8(fl.weight.data)
        obs = L.FP8StaticLinearQuantizer(K, N, wq, ws, fl.bias.data if bias else None, quantize_output=False)
        for _ in range(3):
            obs(make_x(gen, (9, K), torch.float32, zero_row=False))
        mod = L.FP8LinearStatic.from_float(obs)
        x = make_x(gen, (7, K), torch.float32, zero_row=False) * 0.5
        add(f"s{n}", mod, dict(cls="FP8LinearStatic", ctor=dict(in_features=K, out_features=N, use_bias=bias)), x, mod(x))
        # the checkpoint has no `output_scale` (from_float assigned None): a constructor-built module keeps its default
        # 1.0 after loading it, i.e. fake-quantises its output through e4m3 at scale 1 (linear.py:538-540, 562-564)
        fresh = L.FP8LinearStatic(K, N, bias)
        res = fresh.load_state_dict(mod.state_dict(), strict=False)
        cases[-1]["nonstrict_missing"] = list(res.missing_keys)
        arrays[f"s{n}.y_ctor_load"], _ = to_np(fresh(x))
        n += 1
    # decoder-layer-shaped dict: submodule names / classes of models/llama.py:99-106, 206-211
    H, I = 64, 96
    for qc in ({"qkv": "per-tensor", "out": "per-tensor", "fc1": "per-tensor", "fc2": "per-tensor"},
               {"qkv": "per-tensor", "out": "per-token", "fc1": "per-tensor", "fc2": "per-token"}):
        layer = torch.nn.Module()
        layer.self_attn, layer.mlp = torch.nn.Module(), torch.nn.Module()
        for name in ("q_proj", "k_proj", "v_proj"):
            setattr(layer.self_attn, name, L.W8A8BFP32OFP32Linear.from_float(lin(H, H, False), 0.035, act_quant=qc["qkv"]))
        layer.self_attn.o_proj = L.W8A8BFP32OFP32LinearWithQuantScale.from_float(lin(H, H, False), 0.047, act_quant=qc["out"])
        for name in ("gate_proj", "up_proj"):
            setattr(layer.mlp, name, L.W8A8BFP32OFP32Linear.from_float(lin(H, I, False), 0.035, act_quant=qc["fc1"]))
        layer.mlp.down_proj = L.W8A8BFP32OFP32LinearWithQuantScale.from_float(lin(I, H, False), 0.063, act_quant=qc["fc2"])
        add(f"s{n}", layer, dict(cls="decoder_layer", quant_config=qc, hidden=H, intermediate=I))
        n += 1

    OUT_DIR.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT_DIR / "state_golden.npz", **arrays)
    (OUT_DIR / "state_golden.json").write_text(json.dumps({"generator": "oracle/gen_golden_state.py", "torch": torch.__version__,
                                                           "cases": cases}, indent=1))
    print(f"wrote {n} state dicts, {len(arrays)} arrays, {(OUT_DIR / 'state_golden.npz').stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
