"""nn.Linear -> quantized-module conversion of a calibrated, smoothed decoder stack (offline, in place).

What the reference's per-family ``from_float*`` constructors do, as one table:

* INT8 (``models/opt.py:20-29, 88-106, 134-163``; ``models/llama.py:113-133, 222-238, 307-339``;
  ``models/baichuan.py:95-110, 205-223, 250-295``; ``models/mixtral.py:79-93, 107-118, 143-154, 176-221``):
  projections that read a norm's output (q/k/v or the packed ``W_pack``, fc1 / gate / up, Mixtral's w1 / w3) become
  ``W8A8BFP32OFP32Linear`` (``...QKVLinear`` for ``W_pack``); projections that read an activation produced inside the
  block (out / o_proj, fc2 / down / Mixtral's w2 — one scale per expert) become ``...LinearWithQuantScale``; when the
  consumer is per-tensor the preceding norm's weight (and bias) is divided by the input scale so the Linear only
  rounds.  Mixtral's router (``block_sparse_moe.gate``) stays an ``nn.Linear`` (mixtral.py:136-137).
* FP8 (``models/llama.py:137-176, 240-276, 343-353``): every attention / MLP projection becomes ``FP8LinearDynamic``,
  ``FP8LinearStatic`` (from the observer ``calibration.quantize_activations_fp8`` left in its place) or
  ``FP8E5M2Linear``; norms are untouched, static and e5m2 require per-tensor granularity.

Modules are swapped in place; the model's ``state_dict()`` then has the reference's checkpoint schema
(``checkpoint.save_quantized`` writes it).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Sequence, Tuple

import torch
from torch import nn

from ..layers.nn.linear import (FP8E5M2Linear, FP8LinearDynamic, FP8LinearStatic, FP8StaticLinearQuantizer,
                                W8A8BFP32OFP32Linear, W8A8BFP32OFP32LinearWithQuantScale, W8A8BFP32OFP32QKVLinear)
from .smooth import layer_kind

_GRANULARITY_KEYS = ("qkv", "out", "fc1", "fc2")

# family -> [(path of the Linear inside the layer, quant_config key, name of its input scale in the layer's scale dict)]
# Mixtral's experts are expanded per layer (their count is a property of the model).
_INT8_PLAN: Dict[str, List[Tuple[str, str, str]]] = {
    "transformers": [("self_attn.q_proj", "qkv", "attn_input_scale"), ("self_attn.k_proj", "qkv", "attn_input_scale"),
                     ("self_attn.v_proj", "qkv", "attn_input_scale"), ("self_attn.out_proj", "out", "out_input_scale"),
                     ("fc1", "fc1", "fc1_input_scale"), ("fc2", "fc2", "fc2_input_scale")],
    "llama": [("self_attn.q_proj", "qkv", "attn_input_scale"), ("self_attn.k_proj", "qkv", "attn_input_scale"),
              ("self_attn.v_proj", "qkv", "attn_input_scale"), ("self_attn.o_proj", "out", "out_input_scale"),
              ("mlp.gate_proj", "fc1", "gate_input_scale"), ("mlp.up_proj", "fc1", "gate_input_scale"),
              ("mlp.down_proj", "fc2", "down_input_scale")],
    "baichuan": [("self_attn.W_pack", "qkv", "attn_input_scale"), ("self_attn.o_proj", "out", "out_input_scale"),
                 ("mlp.gate_proj", "fc1", "gate_input_scale"), ("mlp.up_proj", "fc1", "gate_input_scale"),
                 ("mlp.down_proj", "fc2", "down_input_scale")],
    "mixtral": [("self_attn.q_proj", "qkv", "attn_input_scale"), ("self_attn.k_proj", "qkv", "attn_input_scale"),
                ("self_attn.v_proj", "qkv", "attn_input_scale"), ("self_attn.o_proj", "out", "out_input_scale")],
}
# family -> [(norm attribute, quant_config key of its consumer, scale folded into it)]
_NORM_FOLDS: Dict[str, List[Tuple[str, str, str]]] = {
    "transformers": [("self_attn_layer_norm", "qkv", "attn_input_scale"), ("final_layer_norm", "fc1", "fc1_input_scale")],
    "llama": [("input_layernorm", "qkv", "attn_input_scale"), ("post_attention_layernorm", "fc1", "gate_input_scale")],
    "baichuan": [("input_layernorm", "qkv", "attn_input_scale"), ("post_attention_layernorm", "fc1", "gate_input_scale")],
    "mixtral": [("input_layernorm", "qkv", "attn_input_scale"), ("post_attention_layernorm", "fc1", "moe_input_scale")],
}


def _granularities(quant_config: Dict[str, str]) -> Dict[str, str]:
    qc = {k: "per-tensor" for k in _GRANULARITY_KEYS}
    qc.update({k: v for k, v in quant_config.items() if k in qc})
    return qc


def _swap(layer: nn.Module, path: str, new: nn.Module) -> None:
    parent, _, leaf = path.rpartition(".")
    setattr(layer.get_submodule(parent) if parent else layer, leaf, new)


def _fold_norm(norm: nn.Module, scale: float) -> None:
    norm.weight.data = norm.weight.data / scale
    if getattr(norm, "bias", None) is not None:
        norm.bias.data = norm.bias.data / scale


def decoder_layers(model: nn.Module, kinds: Sequence[str] = ("transformers", "llama", "baichuan", "mixtral")):
    return [(name, m) for name, m in model.named_modules() if layer_kind(m) in kinds]


def _int8_linear(lin: nn.Linear, path: str, role: str, scale: float, act_quant: str) -> nn.Module:
    if path.endswith("W_pack"):  # Baichuan's packed q|k|v: three equal blocks, one dequant scale each (baichuan.py:84-85)
        third = lin.out_features // 3
        return W8A8BFP32OFP32QKVLinear.from_float(lin, scale, [third] * 3, act_quant=act_quant)
    if role in ("qkv", "fc1"):
        return W8A8BFP32OFP32Linear.from_float(lin, scale, act_quant=act_quant)
    return W8A8BFP32OFP32LinearWithQuantScale.from_float(lin, scale, act_quant=act_quant)


@torch.no_grad()
def quantize_decoder_layers(model: nn.Module, decoder_layer_scales: List[Dict[str, object]],
                            quant_config: Dict[str, str]) -> int:
    """In-place INT8 conversion of every recognised decoder layer (OPT, Llama, Baichuan, Mixtral); returns the count."""
    qc = _granularities(quant_config)
    layers = decoder_layers(model)
    if len(layers) != len(decoder_layer_scales):
        raise ValueError(f"{len(layers)} decoder layers but {len(decoder_layer_scales)} scale dicts")
    for (_name, layer), scales in zip(layers, decoder_layer_scales):
        kind = layer_kind(layer)
        plan = list(_INT8_PLAN[kind])
        expert_scale: Dict[str, float] = {}
        if kind == "mixtral":
            per_expert = scales["down_input_scales"]
            experts = layer.block_sparse_moe.experts
            if len(per_expert) != len(experts):
                raise ValueError(f"{len(experts)} experts but {len(per_expert)} down_input_scales")
            for e in range(len(experts)):
                base = f"block_sparse_moe.experts.{e}"
                plan += [(f"{base}.w1", "fc1", "moe_input_scale"), (f"{base}.w3", "fc1", "moe_input_scale"),
                         (f"{base}.w2", "fc2", f"#{e}")]
                expert_scale[f"#{e}"] = per_expert[e]
        for path, role, scale_name in plan:
            scale = expert_scale[scale_name] if scale_name in expert_scale else scales[scale_name]
            _swap(layer, path, _int8_linear(layer.get_submodule(path), path, role, scale, qc[role]))
        for norm_attr, role, scale_name in _NORM_FOLDS[kind]:
            if qc[role] == "per-tensor":
                _fold_norm(getattr(layer, norm_attr), scales[scale_name])
    return len(layers)


def _projection_paths(layer: nn.Module) -> Iterable[str]:
    """The attention / MLP projections of one decoder layer, router excluded."""
    kind = layer_kind(layer)
    for path, _role, _scale in _INT8_PLAN[kind]:
        yield path
    if kind == "mixtral":
        for e in range(len(layer.block_sparse_moe.experts)):
            for w in ("w1", "w3", "w2"):
                yield f"block_sparse_moe.experts.{e}.{w}"


@torch.no_grad()
def quantize_linears_fp8(model: nn.Module, quant_config: Dict[str, str], reference_compat: bool = True) -> int:
    """In-place FP8 conversion of the decoder layers' projections; returns how many were converted.

    ``quant_config["type"]``: "fp8" / "fp8_e4m3" with ``activation_scheme`` "dynamic" (default) or "static" — the latter
    expects ``calibration.quantize_activations_fp8`` to have run, i.e. the projections already ARE observers — or
    "fp8_e5m2".  ``reference_compat`` is handed to ``FP8LinearDynamic.from_float`` (see there)."""
    qtype = "fp8_e4m3" if quant_config.get("type") == "fp8" else quant_config.get("type")
    scheme = quant_config.get("activation_scheme", "dynamic")
    if qtype not in ("fp8_e4m3", "fp8_e5m2"):
        raise ValueError(f"Unsupported quant type: {qtype}")
    if qtype == "fp8_e5m2" or scheme == "static":  # per-tensor only (models/llama.py:143-147, 155-159)
        qc = _granularities(quant_config)
        if qc["qkv"] != "per-tensor" or qc["out"] != "per-tensor":
            raise AssertionError(f"{qtype} {scheme if qtype == 'fp8_e4m3' else ''} supports per-tensor only".replace("  ", " "))
    make: Callable[[nn.Module], nn.Module]
    if qtype == "fp8_e5m2":
        make = FP8E5M2Linear.from_float
    elif scheme == "static":
        def make(obs):
            if not isinstance(obs, FP8StaticLinearQuantizer) or obs.input_scale is None:
                raise ValueError("static FP8 conversion needs calibrated observers: run quantize_activations_fp8 first")
            return FP8LinearStatic.from_float(obs)
    else:
        def make(lin):
            return FP8LinearDynamic.from_float(lin, reference_compat=reference_compat)
    converted = 0
    for _name, layer in decoder_layers(model):
        for path in _projection_paths(layer):
            _swap(layer, path, make(layer.get_submodule(path)))
            converted += 1
    return converted
