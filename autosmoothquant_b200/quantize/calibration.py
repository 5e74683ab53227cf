"""Calibration statistics for SmoothQuant (reference: ``autosmoothquant/quantize/calibration.py:44-244``).

Same forward-hook logic as the reference; the reference also tokenises a JSON dataset inside these functions
(``load_dataset`` + tokenizer, calibration.py:76-82, 221-227) — here the caller passes an iterable of ready
``input_ids`` batches, so the functions have no dataset / tokenizer dependency.
"""
from __future__ import annotations

import functools
from collections import defaultdict
from typing import Dict, Iterable, List, Tuple

import torch
from torch import nn


@torch.no_grad()
def get_act_scales(model: nn.Module, batches: Iterable[torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Per-input-channel absmax of every nn.Linear's input over the calibration batches (calibration.py:44-88)."""
    model.eval()
    device = next(model.parameters()).device
    act_scales: Dict[str, torch.Tensor] = {}

    def stat_tensor(name, tensor):
        hidden_dim = tensor.shape[-1]
        tensor = tensor.view(-1, hidden_dim).abs().detach()
        coming_max = torch.max(tensor, dim=0)[0].float().cpu()
        act_scales[name] = torch.max(act_scales[name], coming_max) if name in act_scales else coming_max

    def stat_input_hook(m, x, y, name):
        stat_tensor(name, x[0] if isinstance(x, tuple) else x)

    hooks = [m.register_forward_hook(functools.partial(stat_input_hook, name=name))
             for name, m in model.named_modules() if isinstance(m, nn.Linear)]
    try:
        for input_ids in batches:
            model(input_ids.to(device))
    finally:
        for h in hooks:
            h.remove()
    return act_scales


_LAYER_KEYS = {
    # model_type: (prefix, {scale name: (linear name, 'input' | 'output')})   calibration.py:90-184
    "transformers": ("model.decoder.layers", {
        "attn_input_scale": ("self_attn.q_proj", "input"), "q_output_scale": ("self_attn.q_proj", "output"),
        "k_output_scale": ("self_attn.k_proj", "output"), "v_output_scale": ("self_attn.v_proj", "output"),
        "out_input_scale": ("self_attn.out_proj", "input"), "fc1_input_scale": ("fc1", "input"),
        "fc2_input_scale": ("fc2", "input")}),
    "llama": ("model.layers", {
        "attn_input_scale": ("self_attn.q_proj", "input"), "q_output_scale": ("self_attn.q_proj", "output"),
        "k_output_scale": ("self_attn.k_proj", "output"), "v_output_scale": ("self_attn.v_proj", "output"),
        "out_input_scale": ("self_attn.o_proj", "input"), "gate_input_scale": ("mlp.gate_proj", "input"),
        "down_input_scale": ("mlp.down_proj", "input")}),
}


@torch.no_grad()
def get_static_decoder_layer_scales(model: nn.Module, batches: Iterable[torch.Tensor], model_type: str = "transformers",
                                    num_layers: int = None) -> Tuple[List[Dict[str, float]], Dict[str, Dict[str, float]]]:
    """Per-tensor input / output absmax of every nn.Linear, then the per-layer scale dicts (= absmax / 127) the
    quantized model classes consume (calibration.py:186-244 + collect_*_layer_scales :90-184)."""
    if model_type not in _LAYER_KEYS:
        raise ValueError(f"unsupport model type: {model_type}")
    model.eval()
    device = next(model.parameters()).device
    act_dict: Dict[str, Dict[str, float]] = defaultdict(dict)

    def stat_io_hook(m, x, y, name):
        x = x[0] if isinstance(x, tuple) else x
        y = y[0] if isinstance(y, tuple) else y
        for key, t in (("input", x), ("output", y)):
            v = t.detach().abs().max().item()
            act_dict[name][key] = max(act_dict[name][key], v) if key in act_dict[name] else v

    hooks = [m.register_forward_hook(functools.partial(stat_io_hook, name=name))
             for name, m in model.named_modules() if isinstance(m, nn.Linear)]
    try:
        for input_ids in batches:
            model(input_ids.to(device))
    finally:
        for h in hooks:
            h.remove()
    prefix, keys = _LAYER_KEYS[model_type]
    if num_layers is None:
        num_layers = model.config.num_hidden_layers
    layer_scales = []
    for idx in range(num_layers):
        layer_scales.append({scale: act_dict[f"{prefix}.{idx}.{lin}"][io] / 127 for scale, (lin, io) in keys.items()})
    return layer_scales, act_dict
