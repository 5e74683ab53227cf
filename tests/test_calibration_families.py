"""SURVEY 8(f)4 remainder: per-family calibration collectors, static FP8 calibration, converters, checkpoint I/O.

The expectations in tests/golden/calib_golden.{npz,json} are outputs of the reference's own, unmodified
``autosmoothquant/quantize/calibration.py`` run on the tiny named-alike models of ``oracle/tiny_models.py``
(generator: oracle/gen_golden_calib.py).  All CPU, a few seconds.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch
from torch import nn

from autosmoothquant_b200.layers.nn.linear import (FP8E5M2Linear, FP8LinearDynamic, FP8LinearStatic,
                                                   FP8StaticLinearQuantizer, W8A8BFP32OFP32Linear,
                                                   W8A8BFP32OFP32LinearWithQuantScale, W8A8BFP32OFP32QKVLinear)
from autosmoothquant_b200.quantize import (get_act_scales, get_static_decoder_layer_scales, quantize_decoder_layers,
                                           quantize_linears_fp8, smooth_lm)
from autosmoothquant_b200.quantize.calibration import get_layers_to_ignore, quantize_activations_fp8
from autosmoothquant_b200.quantize.checkpoint import load_quantized, save_quantized
from oracle.tiny_models import TinyLM, calibration_batches

GOLDEN = Path(__file__).resolve().parent / "golden"
META = json.loads((GOLDEN / "calib_golden.json").read_text())
ARRAYS = np.load(GOLDEN / "calib_golden.npz")
FAMILIES = ("transformers", "llama", "baichuan", "mixtral")
RTOL = 1e-6  # absmax of fp32 CPU matmul outputs; bit-equal on the machine that wrote the goldens


@pytest.mark.parametrize("kind", FAMILIES)
def test_act_scales_match_reference(kind):
    model = TinyLM(kind).eval()
    got = get_act_scales(model, calibration_batches())
    want = {k[len(kind) + 5:]: ARRAYS[k] for k in ARRAYS.files if k.startswith(f"{kind}.act.")}
    assert set(got) == set(want)
    for name, v in got.items():
        assert v.dtype == torch.float32 and v.device.type == "cpu"
        np.testing.assert_allclose(v.numpy(), want[name], rtol=RTOL, err_msg=name)
    if kind == "mixtral":  # every expert saw every token while calibrating, and the router is back to top-2 afterwards
        e0 = got["model.layers.0.block_sparse_moe.experts.0.w1"]
        assert torch.equal(e0, got["model.layers.0.block_sparse_moe.experts.3.w1"])
        assert torch.equal(e0, got["model.layers.0.block_sparse_moe.gate"])
        assert model.model.layers[0].block_sparse_moe.top_k == 2 == META["families"]["mixtral"]["top_k_after"]


@pytest.mark.parametrize("kind", FAMILIES)
def test_static_layer_scales_match_reference(kind):
    model = TinyLM(kind).eval()
    layer_scales, act_dict = get_static_decoder_layer_scales(model, calibration_batches(), model_type=kind)
    want = META["families"][kind]
    assert set(act_dict) == set(want["act_dict"])
    for name, io in act_dict.items():
        for key in ("input", "output"):
            assert io[key] == pytest.approx(want["act_dict"][name][key], rel=RTOL), (name, key)
    assert len(layer_scales) == len(want["layer_scales"]) == 2
    for got, exp in zip(layer_scales, want["layer_scales"]):
        assert set(got) == set(exp)
        for key, v in exp.items():
            if isinstance(v, list):
                assert len(got[key]) == 4 and got[key] == pytest.approx(v, rel=RTOL)
            else:
                assert got[key] == pytest.approx(v, rel=RTOL), key


def test_unrouted_expert_has_no_static_scale():
    """The static pass keeps the model's own top-k (the reference raises it only for the smoothing statistics): an
    expert that no calibration token reached has no entry and the collector raises, as collect_mixtral_layer_scales does."""
    model = TinyLM("mixtral").eval()
    with torch.no_grad():
        for layer in model.model.layers:
            layer.block_sparse_moe.gate.weight[3].fill_(-50.0)  # expert 3 never wins
            layer.block_sparse_moe.gate.weight[3, 0] = 0.0
    with pytest.raises(KeyError):
        get_static_decoder_layer_scales(model, calibration_batches(n=1, seq=4), model_type="mixtral")
    with pytest.raises(ValueError):
        get_static_decoder_layer_scales(model, [], model_type="gpt2")


def test_layers_to_ignore_matches_reference():
    model = TinyLM("llama")
    assert sorted(get_layers_to_ignore(model, ["re:.*lm_head"])) == META["ignore"]["re:.*lm_head"] == ["lm_head"]
    got = sorted(get_layers_to_ignore(model, ["model.layers.0.mlp.down_proj", "re:layers\\.1\\.self_attn"]))
    assert got == META["ignore"]["exact+regex"] and len(got) == 5


def test_fp8_static_calibration_matches_reference():
    model = TinyLM("llama").eval()
    n = quantize_activations_fp8(model, calibration_batches(), ["re:.*lm_head"])
    assert n == len(META["fp8_static"]) == 14 and isinstance(model.lm_head, nn.Linear) and META["lm_head_is_linear"]
    observers = {name: m for name, m in model.named_modules() if isinstance(m, FP8StaticLinearQuantizer)}
    assert set(observers) == set(META["fp8_static"])
    assert quantize_linears_fp8(model, {"type": "fp8", "activation_scheme": "static", "qkv": "per-tensor", "out": "per-tensor",
                                        "fc1": "per-tensor", "fc2": "per-tensor"}) == 14
    for name, exp in META["fp8_static"].items():
        mod = model.get_submodule(name)
        assert isinstance(mod, FP8LinearStatic)
        np.testing.assert_array_equal(mod.weight.view(torch.uint8).numpy(), ARRAYS[f"fp8static.{name}.weight"])
        assert float(mod.weight_scale) == pytest.approx(exp["weight_scale"], rel=RTOL)
        assert float(mod.input_scale) == pytest.approx(exp["input_scale"], rel=RTOL)
        assert exp["output_scale"] is None and float(mod.output_scale) == 0.0  # falsy either way: no output fake-quant
        for buf in (mod.weight_scale, mod.input_scale):
            assert buf.dtype == torch.float32 and buf.device.type == "cpu" and buf.dim() == 0


def test_fp8_static_conversion_needs_observers():
    with pytest.raises(ValueError):
        quantize_linears_fp8(TinyLM("llama"), {"type": "fp8_e4m3", "activation_scheme": "static"})
    with pytest.raises(AssertionError):  # static / e5m2 are per-tensor only (models/llama.py:143-147, 155-159)
        quantize_linears_fp8(TinyLM("llama"), {"type": "fp8_e5m2", "qkv": "per-token"})
    with pytest.raises(ValueError):
        quantize_linears_fp8(TinyLM("llama"), {"type": "int8"})


@pytest.mark.parametrize("qtype,cls", [("fp8", FP8LinearDynamic), ("fp8_e5m2", FP8E5M2Linear)])
def test_fp8_dynamic_and_e5m2_conversion(qtype, cls):
    model = TinyLM("llama").eval()
    floats = {n: m.weight.detach().clone() for n, m in model.named_modules() if isinstance(m, nn.Linear)}
    assert quantize_linears_fp8(model, {"type": qtype}) == 14
    assert isinstance(model.lm_head, nn.Linear)
    for name, w in floats.items():
        if name == "lm_head":
            continue
        mod = model.get_submodule(name)
        assert isinstance(mod, cls)
        if cls is FP8LinearDynamic:
            assert mod.weight.dtype == torch.float8_e4m3fn
            back = mod.weight.float() * float(mod.weight_scale)
            assert float((back - w).abs().max()) <= float(w.abs().max()) * 2 ** -4  # e4m3: 3 mantissa bits
            assert mod.act_quant is False  # the reference converter's positional quirk, kept by default (linear.py:444-446)
        else:
            assert mod.weight.dtype == torch.float8_e5m2
    if cls is FP8LinearDynamic:
        model = TinyLM("llama")
        quantize_linears_fp8(model, {"type": "fp8_e4m3"}, reference_compat=False)
        assert model.model.layers[1].mlp.down_proj.act_quant == "per-token"


@pytest.mark.parametrize("kind", FAMILIES)
@pytest.mark.parametrize("granularity", ["per-tensor", "per-token"])
def test_int8_conversion_of_every_family(kind, granularity, tmp_path):
    model = TinyLM(kind).eval()
    batches = calibration_batches()
    assert smooth_lm(model, get_act_scales(model, batches), alpha=0.5) == 2
    layer_scales, _ = get_static_decoder_layer_scales(model, batches, model_type=kind)
    qc = {k: granularity for k in ("qkv", "out", "fc1", "fc2")}
    layer0 = (model.model.decoder if kind == "transformers" else model.model).layers[0]
    norm_names = ("self_attn_layer_norm", "final_layer_norm") if kind == "transformers" else ("input_layernorm", "post_attention_layernorm")
    norms_before = [getattr(layer0, n).weight.detach().clone() for n in norm_names]
    assert quantize_decoder_layers(model, layer_scales, qc) == 2
    s = layer_scales[0]
    second = {"transformers": "fc1_input_scale", "mixtral": "moe_input_scale"}.get(kind, "gate_input_scale")
    for n, before, scale in zip(norm_names, norms_before, (s["attn_input_scale"], s[second])):
        after = getattr(layer0, n).weight.detach()
        torch.testing.assert_close(after, before / scale if granularity == "per-tensor" else before)
    attn = layer0.self_attn
    if kind == "baichuan":
        assert isinstance(attn.W_pack, W8A8BFP32OFP32QKVLinear) and attn.W_pack.qkv_size == [32, 32, 32]
        for nm in ("q_dequant_scale", "k_dequant_scale", "v_dequant_scale"):
            assert getattr(attn.W_pack, nm).dim() == 0
    else:
        assert all(isinstance(getattr(attn, p), W8A8BFP32OFP32Linear) for p in ("q_proj", "k_proj", "v_proj"))
    out_proj = attn.out_proj if kind == "transformers" else attn.o_proj
    assert isinstance(out_proj, W8A8BFP32OFP32LinearWithQuantScale) and out_proj.act_quant == granularity
    per_tensor = granularity == "per-tensor"  # the quant_scale buffer exists for per-tensor only (linear.py:262-266)
    assert hasattr(out_proj, "quant_scale") == per_tensor
    if per_tensor:
        assert float(out_proj.quant_scale) == pytest.approx(s["out_input_scale"])
    if kind == "mixtral":
        moe = layer0.block_sparse_moe
        assert isinstance(moe.gate, nn.Linear)  # the router is not quantized (models/mixtral.py:136-137)
        for e, expert in enumerate(moe.experts):
            assert isinstance(expert.w1, W8A8BFP32OFP32Linear) and isinstance(expert.w3, W8A8BFP32OFP32Linear)
            assert isinstance(expert.w2, W8A8BFP32OFP32LinearWithQuantScale)
            if per_tensor:
                assert float(expert.w2.quant_scale) == pytest.approx(s["down_input_scales"][e])
    elif kind == "transformers":
        assert isinstance(layer0.fc1, W8A8BFP32OFP32Linear) and isinstance(layer0.fc2, W8A8BFP32OFP32LinearWithQuantScale)
    else:
        assert isinstance(layer0.mlp.up_proj, W8A8BFP32OFP32Linear)
        assert isinstance(layer0.mlp.down_proj, W8A8BFP32OFP32LinearWithQuantScale)
    with pytest.raises(ValueError):
        quantize_decoder_layers(TinyLM(kind), layer_scales[:1], qc)
    # ---- checkpoint round trip in the reference's on-disk format (examples/smoothquant_model.py:96-99)
    out = save_quantized(model, tmp_path / kind, dict(qc, type="int8"))
    state, cfg = load_quantized(out)
    assert cfg == dict(qc, type="int8")
    want = model.state_dict()
    assert set(state) == set(want)
    for k, v in want.items():
        assert state[k].dtype == v.dtype and torch.equal(state[k], v.cpu()), k
    fresh = TinyLM(kind)
    quantize_decoder_layers(fresh, layer_scales, qc)  # same module structure, different (seed-0 float) contents
    with torch.no_grad():
        for p in fresh.parameters():
            p.zero_()
    fresh.load_state_dict(state, strict=True)
    for k, v in fresh.state_dict().items():
        assert torch.equal(v, want[k]), k


def test_fp8_checkpoint_round_trip(tmp_path):
    model = TinyLM("llama").eval()
    quantize_activations_fp8(model, calibration_batches())
    cfg_in = {"type": "fp8", "activation_scheme": "static", "qkv": "per-tensor", "out": "per-tensor", "fc1": "per-tensor",
              "fc2": "per-tensor"}
    quantize_linears_fp8(model, cfg_in)
    out = save_quantized(model, tmp_path / "fp8", cfg_in)
    state, cfg = load_quantized(out)
    assert cfg["type"] == "fp8_e4m3"  # normalised before it is written (smoothquant_model.py:69-70)
    w = state["model.layers.0.self_attn.q_proj.weight"]
    assert w.dtype == torch.float8_e4m3fn
    assert torch.equal(w.view(torch.uint8), model.model.layers[0].self_attn.q_proj.weight.view(torch.uint8))
    assert state["model.layers.1.mlp.down_proj.input_scale"].dtype == torch.float32
