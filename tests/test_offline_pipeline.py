"""Offline pipeline (BASELINE config 0: OPT-125M per-tensor INT8 calibrate + quantize on CPU via torch).

* smooth_ln_fcs against golden vectors produced by the reference's own smooth.py (oracle/gen_golden_smooth.py);
* the whole pipeline on a random-init OPT-125M (true shapes, HF implementation of the installed transformers):
  calibration hooks -> smoothing (output-preserving) -> static per-layer scales -> module conversion; the state
  dict must have the reference's checkpoint schema and the int8 weights must reproduce the smoothed weights to
  half a quantisation step.  No GPU involved: this is the plumbing configuration.
"""
from pathlib import Path

import numpy as np
import pytest
import torch
from torch import nn

from autosmoothquant_b200.quantize import (get_act_scales, get_static_decoder_layer_scales, quantize_decoder_layers,
                                           smooth_lm, smooth_ln_fcs)

GOLDEN = Path(__file__).resolve().parent / "golden" / "smooth_golden.npz"


class _RMSNorm(nn.Module):  # weight-only norm, as LlamaRMSNorm
    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))


@pytest.mark.parametrize("tag,model_type", [("opt_qkv", "transformers"), ("opt_fc1", "transformers"), ("llama_gateup", "llama")])
def test_smooth_ln_fcs_matches_reference_golden(tag, model_type):
    z = np.load(GOLDEN)
    n = z[f"{tag}.ln_weight_in"].shape[0]
    ln = nn.LayerNorm(n) if model_type == "transformers" else _RMSNorm(n)
    fcs = []
    with torch.no_grad():
        ln.weight.copy_(torch.from_numpy(z[f"{tag}.ln_weight_in"]))
        if model_type == "transformers":
            ln.bias.copy_(torch.from_numpy(z[f"{tag}.ln_bias_in"]))
        i = 0
        while f"{tag}.fc{i}_in" in z:
            w = torch.from_numpy(z[f"{tag}.fc{i}_in"])
            fc = nn.Linear(w.shape[1], w.shape[0])
            fc.weight.copy_(w)
            fcs.append(fc)
            i += 1
    smooth_ln_fcs(ln, fcs if len(fcs) > 1 else fcs[0], torch.from_numpy(z[f"{tag}.act_scales"]), model_type, float(z[f"{tag}.alpha"]))
    np.testing.assert_array_equal(ln.weight.detach().numpy(), z[f"{tag}.ln_weight_out"])
    if model_type == "transformers":
        np.testing.assert_array_equal(ln.bias.detach().numpy(), z[f"{tag}.ln_bias_out"])
    for i, fc in enumerate(fcs):
        np.testing.assert_array_equal(fc.weight.detach().numpy(), z[f"{tag}.fc{i}_out"])


@pytest.fixture(scope="module")
def opt125m():
    from transformers import OPTConfig, OPTForCausalLM

    cfg = OPTConfig(vocab_size=50272, hidden_size=768, ffn_dim=3072, num_hidden_layers=12, num_attention_heads=12,
                    max_position_embeddings=2048, word_embed_proj_dim=768)
    torch.manual_seed(0)
    return OPTForCausalLM(cfg).eval()


def test_opt125m_calibrate_smooth_quantize_on_cpu(opt125m):
    model = opt125m
    g = torch.Generator().manual_seed(1)
    batches = [torch.randint(0, 50272, (1, 128), generator=g) for _ in range(8)]  # 8 synthetic samples x 128 tokens
    with torch.no_grad():
        before = model(batches[0]).logits
    act_scales = get_act_scales(model, batches)
    assert len(act_scales) == 12 * 6 + 1  # six nn.Linear per decoder layer + lm_head
    q0 = act_scales["model.decoder.layers.0.self_attn.q_proj"]
    assert q0.shape == (768,) and q0.dtype == torch.float32 and float(q0.min()) > 0
    # k_proj / v_proj see the same input as q_proj
    assert torch.equal(q0, act_scales["model.decoder.layers.0.self_attn.k_proj"])
    assert smooth_lm(model, act_scales, alpha=0.5) == 12
    with torch.no_grad():
        after = model(batches[0]).logits
    # smoothing is an exact re-parametrisation: (x / s) @ (W * s)^T == x @ W^T up to fp32 rounding
    assert float((after - before).abs().max()) <= 2e-3 * float(before.abs().max())
    layer_scales, act_dict = get_static_decoder_layer_scales(model, batches, model_type="transformers")
    assert len(layer_scales) == 12 and set(layer_scales[0]) == {
        "attn_input_scale", "q_output_scale", "k_output_scale", "v_output_scale", "out_input_scale", "fc1_input_scale",
        "fc2_input_scale"}
    assert layer_scales[3]["fc1_input_scale"] == act_dict["model.decoder.layers.3.fc1"]["input"] / 127
    smoothed = {n: m.weight.detach().clone() for n, m in model.named_modules() if isinstance(m, nn.Linear)}
    ln_w = model.model.decoder.layers[0].self_attn_layer_norm.weight.detach().clone()
    assert quantize_decoder_layers(model, layer_scales, {"qkv": "per-tensor", "out": "per-tensor", "fc1": "per-tensor",
                                                         "fc2": "per-tensor"}) == 12
    layer0 = model.model.decoder.layers[0]
    # norm folding (models/opt.py:20-29): the LayerNorm now emits "int8 units"
    torch.testing.assert_close(layer0.self_attn_layer_norm.weight.detach(), ln_w / layer_scales[0]["attn_input_scale"])
    sd = model.state_dict()
    pre = "model.decoder.layers.0."
    for proj, has_qs in (("self_attn.q_proj", False), ("self_attn.out_proj", True), ("fc1", False), ("fc2", True)):
        w = sd[pre + proj + ".weight"]
        assert w.dtype == torch.int8 and int(w.abs().max()) == 127
        assert sd[pre + proj + ".bias"].dtype == torch.float32
        ds = sd[pre + proj + ".dequant_scale"]
        assert ds.dtype == torch.float32 and ds.dim() == 0 and ds.device.type == "cpu"
        assert ((pre + proj + ".quant_scale") in sd) == has_qs
        # dequant_scale = input_scale * weight_scale (linear.py:120-123); int8 weights reproduce the smoothed weights
        key = {"self_attn.q_proj": "attn_input_scale", "self_attn.out_proj": "out_input_scale", "fc1": "fc1_input_scale",
               "fc2": "fc2_input_scale"}[proj]
        w_scale = float(ds) / layer_scales[0][key]
        ref_w = smoothed["model.decoder.layers.0." + proj]
        assert float((w.float() * w_scale - ref_w).abs().max()) <= 0.5 * w_scale * 1.001


def test_hf_llama_calibrate_smooth_quantize_save_on_cpu(tmp_path):
    """The same chain on the INSTALLED transformers' LlamaForCausalLM (random-init, small): the pipeline finds HF's decoder
    layers by structure, the static scales carry the reference's names (collect_llama_layer_scales), per-tensor qkv / fc1
    fold their input scale into the RMSNorm weights, out / fc2 stay per-token (BASELINE config 3's granularities), and the
    saved directory is the reference's on-disk pair with its state-dict keys."""
    from transformers import LlamaConfig, LlamaForCausalLM

    from autosmoothquant_b200.layers.nn.linear import W8A8BFP32OFP32Linear, W8A8BFP32OFP32LinearWithQuantScale
    from autosmoothquant_b200.quantize import load_quantized, save_quantized

    cfg = LlamaConfig(vocab_size=320, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=128)
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg).eval()
    g = torch.Generator().manual_seed(1)
    batches = [torch.randint(0, 320, (1, 48), generator=g) for _ in range(4)]
    with torch.no_grad():
        before = model(batches[0]).logits
    act_scales = get_act_scales(model, batches)
    assert "model.layers.1.mlp.down_proj" in act_scales and act_scales["model.layers.0.self_attn.q_proj"].shape == (64,)
    assert smooth_lm(model, act_scales, alpha=0.5) == 2
    with torch.no_grad():
        after = model(batches[0]).logits
    assert float((after - before).abs().max()) <= 2e-3 * float(before.abs().max())  # an exact re-parametrisation
    layer_scales, _ = get_static_decoder_layer_scales(model, batches, model_type="llama")
    assert set(layer_scales[0]) == {"attn_input_scale", "q_output_scale", "k_output_scale", "v_output_scale", "out_input_scale",
                                    "gate_input_scale", "down_input_scale"}
    qc = {"qkv": "per-tensor", "out": "per-token", "fc1": "per-tensor", "fc2": "per-token", "type": "int8"}
    ln0 = model.model.layers[0].input_layernorm.weight.detach().clone()
    assert quantize_decoder_layers(model, layer_scales, qc) == 2
    layer0 = model.model.layers[0]
    torch.testing.assert_close(layer0.input_layernorm.weight.detach(), ln0 / layer_scales[0]["attn_input_scale"])
    assert isinstance(layer0.self_attn.k_proj, W8A8BFP32OFP32Linear) and layer0.self_attn.k_proj.weight.shape == (32, 64)
    assert isinstance(layer0.mlp.down_proj, W8A8BFP32OFP32LinearWithQuantScale) and layer0.mlp.down_proj.act_quant == "per-token"
    assert not hasattr(layer0.mlp.down_proj, "quant_scale")  # per-token modules carry no static input scale (linear.py:252-256)
    assert isinstance(model.lm_head, nn.Linear)  # not quantized, as in the reference
    out = save_quantized(model, tmp_path / "llama-smoothquant-int8", qc)
    state, cfg2 = load_quantized(out)
    assert cfg2 == qc
    assert state["model.layers.1.self_attn.o_proj.weight"].dtype == torch.int8
    assert state["model.layers.1.mlp.gate_proj.dequant_scale"].dim() == 0
    assert "model.layers.0.self_attn.o_proj.quant_scale" not in state and "model.embed_tokens.weight" in state
