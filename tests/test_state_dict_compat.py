"""N1 — what the reference WRITES loads into the B200 modules and computes the same thing.

tests/golden/state_golden.* holds `state_dict()`s of the reference's own modules (oracle/gen_golden_state.py: built by
its from_float converters / constructors, unmodified code on CPU), one input and the reference's output each, plus
decoder-layer-shaped dicts keyed as models/llama.py:99-106, 206-211 name the projections.

CPU (not gpu): load_state_dict(strict=True) into modules built by the same constructor arguments, identical key set /
dtypes / shapes / values, scales stay 0-dim fp32 on the host, save -> load round trip through the on-disk format
(safetensors + quant_config.json, FP8 weights included).
GPU: the loaded module's forward equals the reference's recorded output (INT8: bit for bit in the exact-division mode
the CPU reference computes in; FP8: the tensor-core tolerance of test_gpu_parity.test_golden_fp8_linear).
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch
from torch import nn

from autosmoothquant_b200.layers.nn import linear as NN
from autosmoothquant_b200.quantize import checkpoint

GOLD = Path(__file__).resolve().parent / "golden"
TORCH_DT = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16, "int8": torch.int8}


@pytest.fixture(scope="module")
def state_cases():
    meta = json.loads((GOLD / "state_golden.json").read_text())
    arrays = np.load(GOLD / "state_golden.npz")
    cases = []
    for c in meta["cases"]:
        sd = {}
        for k in c["keys"]:
            a = torch.from_numpy(arrays[f"{c['id']}.sd.{k}"])
            dt = c["dtypes"][k]
            sd[k] = a.view(torch.float8_e4m3fn) if dt == "float8_e4m3fn" else a.to(TORCH_DT[dt])
            assert list(sd[k].shape) == c["shapes"][k]
        entry = dict(c, sd=sd)
        for extra in ("x", "y", "y_ctor_load"):
            if f"{c['id']}.{extra}" in arrays.files:
                entry[extra] = arrays[f"{c['id']}.{extra}"]
        cases.append(entry)
    return cases


def build_module(c) -> nn.Module:
    """The B200 module a model constructor (or, for the converter's quirky FP8 product, from_float) would create."""
    cls, ctor = c["cls"], c.get("ctor", {})
    if cls == "decoder_layer":
        qc, H, I = c["quant_config"], c["hidden"], c["intermediate"]
        layer = nn.Module()
        layer.self_attn, layer.mlp = nn.Module(), nn.Module()
        for name in ("q_proj", "k_proj", "v_proj"):  # models/llama.py:99-106
            setattr(layer.self_attn, name, NN.W8A8BFP32OFP32Linear(H, H, act_quant=qc["qkv"]))
        layer.self_attn.o_proj = NN.W8A8BFP32OFP32LinearWithQuantScale(H, H, act_quant=qc["out"])
        for name in ("gate_proj", "up_proj"):  # models/llama.py:206-211
            setattr(layer.mlp, name, NN.W8A8BFP32OFP32Linear(H, I, act_quant=qc["fc1"]))
        layer.mlp.down_proj = NN.W8A8BFP32OFP32LinearWithQuantScale(I, H, act_quant=qc["fc2"])
        return layer
    if cls == "FP8LinearDynamic" and c["how"] == "from_float":
        has_bias = "bias" in c["keys"]
        return NN.FP8LinearDynamic.from_float(nn.Linear(ctor["in_features"], ctor["out_features"], bias=has_bias))  # reference_compat
    if cls == "W8A8BFP32OFP32QKVLinear":
        return NN.W8A8BFP32OFP32QKVLinear(ctor["qkv_size"], ctor["in_features"], ctor["out_features"], ctor["use_bias"], ctor["act_quant"])
    if cls == "FP8LinearDynamic":
        return NN.FP8LinearDynamic(ctor["in_features"], ctor["out_features"], ctor["act_quant"], ctor["use_bias"])
    if cls == "FP8LinearStatic":
        return NN.FP8LinearStatic(ctor["in_features"], ctor["out_features"], ctor["use_bias"])
    return getattr(NN, cls)(ctor["in_features"], ctor["out_features"], ctor["use_bias"], ctor["act_quant"])


def load(c) -> nn.Module:
    mod = build_module(c)
    if c["cls"] == "FP8LinearStatic":
        # the reference's converter leaves `output_scale` out of the dict (it assigns None, linear.py:580): a strict load
        # fails for the reference's own constructor-built module too; non-strict reports exactly that key
        res = mod.load_state_dict(c["sd"], strict=False)
        assert list(res.missing_keys) == c["nonstrict_missing"] == ["output_scale"] and not res.unexpected_keys
    else:
        mod.load_state_dict(c["sd"], strict=True)
    return mod


def test_reference_state_dicts_load_with_identical_schema(state_cases):
    assert len(state_cases) >= 18
    for c in state_cases:
        mod = load(c)
        ours = mod.state_dict()
        extra = {"output_scale"} if c["cls"] == "FP8LinearStatic" else set()
        assert set(ours) - extra == set(c["sd"]), c["id"]
        for k, want in c["sd"].items():
            got = ours[k]
            assert got.dtype == want.dtype and got.shape == want.shape, (c["id"], k, got.dtype, want.dtype)
            assert torch.equal(got.view(torch.uint8) if got.dtype == torch.float8_e4m3fn else got,
                               want.view(torch.uint8) if want.dtype == torch.float8_e4m3fn else want), (c["id"], k)
            if k.endswith("_scale"):  # passed to the kernel by value: 0-dim fp32 on the host (linear.py:68-72)
                assert got.dim() == 0 and got.dtype == torch.float32 and got.device.type == "cpu"


def test_converter_quirk_of_fp8_dynamic_is_reproduced(state_cases):
    """reference_compat (default): the module `from_float` returns carries use_bias in the act_quant slot — per-tensor
    branch, bias stored but unused — exactly like the reference's (linear.py:444-451)."""
    for c in (c for c in state_cases if c["cls"] == "FP8LinearDynamic" and c["how"] == "from_float"):
        mod = load(c)
        assert mod.act_quant == c["ctor"]["act_quant"] and isinstance(mod.act_quant, bool)
        assert mod.use_bias is False and ("bias" in mod.state_dict()) == ("bias" in c["keys"])
    fixed = NN.FP8LinearDynamic.from_float(nn.Linear(16, 8, bias=True), reference_compat=False)
    assert fixed.act_quant == "per-token" and fixed.use_bias is True


def test_checkpoint_round_trip_on_disk(state_cases, tmp_path):
    """save -> load through the reference's on-disk pair (safetensors + quant_config.json), INT8 and FP8 tensors."""
    from autosmoothquant_b200.layers.functional.quantization import dtype_byte_size

    assert dtype_byte_size(torch.float8_e4m3fn) == 1 and dtype_byte_size(torch.bfloat16) == 2 and dtype_byte_size(torch.bool) == 1 / 8
    import transformers.modeling_utils as mu

    assert mu.dtype_byte_size(torch.float8_e4m3fn) == 1  # the patch the reference installs at import (quantization.py:126-136)
    for c in state_cases:
        mod = load(c)
        qc = dict(c.get("quant_config", {"qkv": "per-tensor"}), type="fp8" if c["cls"].startswith("FP8") else "int8")
        out = checkpoint.save_quantized(mod, tmp_path / c["id"], qc)
        state, qc2 = checkpoint.load_quantized(out)
        assert qc2["type"] == ("fp8_e4m3" if c["cls"].startswith("FP8") else "int8")
        fresh = build_module(c)
        fresh.load_state_dict(state, strict=True)  # our own files always carry every buffer
        for k, v in mod.state_dict().items():
            w = fresh.state_dict()[k]
            assert w.dtype == v.dtype and torch.equal(w.view(torch.uint8) if w.dtype == torch.float8_e4m3fn else w,
                                                      v.view(torch.uint8) if v.dtype == torch.float8_e4m3fn else v), (c["id"], k)
        weights = (out / checkpoint.WEIGHTS_NAME).stat().st_size
        payload = sum(v.numel() * v.element_size() for v in mod.state_dict().values())
        assert payload <= weights <= payload + 16384  # one byte per 8-bit weight: nothing was widened on the way


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_loaded_reference_modules_compute_the_reference_outputs(state_cases):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from autosmoothquant_b200 import _lib as L

    dev = torch.device("cuda:0")
    prev = L.set_div_mode(L.DIV_EXACT)  # the goldens were computed by the reference's Python on CPU (true division)
    try:
        checked = 0
        for c in (c for c in state_cases if "x" in c):
            xdt = TORCH_DT[c["x_dtype"]]
            x = torch.from_numpy(c["x"]).to(xdt).to(dev)
            mod = load(c)
            mod = mod.to(dev) if not c["cls"].startswith("FP8") else mod._apply(lambda t: t.to(dev))
            if c["cls"] == "FP8LinearStatic":
                # after the non-strict load the constructor default output_scale = 1.0 is in force, as in the reference
                want = c["y_ctor_load"]
                y = mod(x).float().cpu().numpy()
                scale = np.abs(want).max()
                assert np.mean(y != want) < 0.02 and np.abs(y - want).max() <= 0.13 * scale, c["id"]  # e4m3-grid neighbours
                mod.output_scale = torch.tensor(0.0)  # the converter's own module: no output fake-quantisation
            y = mod(x)
            assert y.dtype == xdt and tuple(y.shape) == c["y"].shape
            got, want = y.float().cpu().numpy(), c["y"]
            if c["cls"].startswith("FP8"):
                np.testing.assert_allclose(got, want, rtol=0, atol=2e-5 * np.abs(want).max(), err_msg=c["id"])
            else:
                np.testing.assert_array_equal(got, want, err_msg=f"{c['id']} {c['cls']}")
            checked += 1
        assert checked >= 16
        # decoder-layer-shaped checkpoints: every projection is live after the load
        for c in (c for c in state_cases if c["cls"] == "decoder_layer"):
            layer = load(c).to(dev)
            h = (torch.randn(5, c["hidden"], generator=torch.Generator().manual_seed(1)) * 30).to(torch.bfloat16).to(dev)
            q = layer.self_attn.q_proj(h)
            o = layer.self_attn.o_proj(q)
            d = layer.mlp.down_proj(torch.nn.functional.silu(layer.mlp.gate_proj(h)) * layer.mlp.up_proj(h))
            torch.cuda.synchronize()
            assert torch.isfinite(o).all() and torch.isfinite(d).all() and float(d.abs().max()) > 0
    finally:
        L.set_div_mode(prev)
