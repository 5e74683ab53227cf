"""nn.Linear -> W8A8 module conversion of a calibrated, smoothed decoder stack.

What the reference's ``Int8OPTDecoderLayer.from_float`` / ``Int8LlamaDecoderLayer.from_float`` do
(``autosmoothquant/models/opt.py:20-29, 88-106, 134-163``; ``models/llama.py:27-37, 326-339``): q/k/v and
fc1 / gate / up become ``W8A8BFP32OFP32Linear``, out / fc2 / down ``W8A8BFP32OFP32LinearWithQuantScale``, and
when the consumer is per-tensor the preceding norm's weight (and bias) is divided by its input scale so the
Linear only rounds.  Modules are swapped in place; the state dict then has the reference's checkpoint schema.
"""
from __future__ import annotations

from typing import Dict, List

import torch
from torch import nn

from ..layers.nn.linear import W8A8BFP32OFP32Linear, W8A8BFP32OFP32LinearWithQuantScale
from .smooth import layer_kind


def _fold_norm(norm: nn.Module, scale: float) -> None:
    norm.weight.data = norm.weight.data / scale
    if getattr(norm, "bias", None) is not None:
        norm.bias.data = norm.bias.data / scale


@torch.no_grad()
def quantize_decoder_layers(model: nn.Module, decoder_layer_scales: List[Dict[str, float]],
                            quant_config: Dict[str, str]) -> int:
    """In-place INT8 conversion of every recognised decoder layer (OPT and Llama families); returns the count."""
    qc = {"qkv": "per-tensor", "out": "per-tensor", "fc1": "per-tensor", "fc2": "per-tensor"}
    qc.update({k: v for k, v in quant_config.items() if k in qc})
    layers = [(name, m) for name, m in model.named_modules() if layer_kind(m) in ("transformers", "llama")]
    if len(layers) != len(decoder_layer_scales):
        raise ValueError(f"{len(layers)} decoder layers but {len(decoder_layer_scales)} scale dicts")
    for (name, layer), s in zip(layers, decoder_layer_scales):
        attn = layer.self_attn
        for proj in ("q_proj", "k_proj", "v_proj"):
            setattr(attn, proj, W8A8BFP32OFP32Linear.from_float(getattr(attn, proj), s["attn_input_scale"], act_quant=qc["qkv"]))
        if layer_kind(layer) == "transformers":
            attn.out_proj = W8A8BFP32OFP32LinearWithQuantScale.from_float(attn.out_proj, s["out_input_scale"], act_quant=qc["out"])
            layer.fc1 = W8A8BFP32OFP32Linear.from_float(layer.fc1, s["fc1_input_scale"], act_quant=qc["fc1"])
            layer.fc2 = W8A8BFP32OFP32LinearWithQuantScale.from_float(layer.fc2, s["fc2_input_scale"], act_quant=qc["fc2"])
            if qc["qkv"] == "per-tensor":
                _fold_norm(layer.self_attn_layer_norm, s["attn_input_scale"])
            if qc["fc1"] == "per-tensor":
                _fold_norm(layer.final_layer_norm, s["fc1_input_scale"])
        else:
            attn.o_proj = W8A8BFP32OFP32LinearWithQuantScale.from_float(attn.o_proj, s["out_input_scale"], act_quant=qc["out"])
            mlp = layer.mlp
            mlp.gate_proj = W8A8BFP32OFP32Linear.from_float(mlp.gate_proj, s["gate_input_scale"], act_quant=qc["fc1"])
            mlp.up_proj = W8A8BFP32OFP32Linear.from_float(mlp.up_proj, s["gate_input_scale"], act_quant=qc["fc1"])
            mlp.down_proj = W8A8BFP32OFP32LinearWithQuantScale.from_float(mlp.down_proj, s["down_input_scale"], act_quant=qc["fc2"])
            if qc["qkv"] == "per-tensor":
                _fold_norm(layer.input_layernorm, s["attn_input_scale"])
            if qc["fc1"] == "per-tensor":
                _fold_norm(layer.post_attention_layernorm, s["gate_input_scale"])
    return len(layers)
