"""Activation smoothing (SmoothQuant): migrate quantisation difficulty from activations to weights.

Behaviour of the reference's ``autosmoothquant/quantize/smooth.py:10-93``: for a norm followed by linears that
share its output, every input channel k gets ``s[k] = act_absmax[k]**alpha / weight_absmax[k]**(1-alpha)`` (both
clamped at 1e-5), the norm's weight (and bias, for LayerNorm) is divided by s and the linears' weight columns are
multiplied by s — an exact re-parametrisation.  The reference dispatches on HF 4.42 classes; here a decoder layer
is recognised by its attribute structure and described by a small table, so the same code serves the installed
transformers and the bench's own stack.  Offline, not timed.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple, Union

import torch
from torch import nn

_FLOOR = 1e-5


@torch.no_grad()
def smooth_ln_fcs(ln: nn.Module, fcs: Union[nn.Linear, List[nn.Linear]], act_scales: torch.Tensor,
                  model_type: str = "transformers", alpha: float = 0.5) -> None:
    """One (norm, linears) group.  ``model_type == "transformers"`` = LayerNorm with bias (OPT); llama / mixtral /
    baichuan norms are weight-only RMSNorms."""
    linears: Sequence[nn.Linear] = fcs if isinstance(fcs, list) else [fcs]
    width = ln.weight.numel()
    for fc in linears:
        if not isinstance(fc, nn.Linear) or fc.in_features != width or act_scales.numel() != width:
            raise AssertionError("smooth_ln_fcs: every linear must consume the norm's output channels")
    has_bias = model_type == "transformers"
    if has_bias and not isinstance(ln, nn.LayerNorm):
        raise AssertionError("model_type 'transformers' expects nn.LayerNorm")
    ref = linears[0].weight
    act = act_scales.to(device=ref.device, dtype=ref.dtype)
    # per input channel: the largest |w| over all output rows of all linears of the group
    w_absmax = torch.cat([fc.weight.abs().max(dim=0, keepdim=True)[0] for fc in linears], dim=0).max(dim=0)[0].clamp(min=_FLOOR)
    s = (act.pow(alpha) / w_absmax.pow(1 - alpha)).clamp(min=_FLOOR).to(ref.device).to(ref.dtype)
    ln.weight.div_(s)
    if has_bias:
        ln.bias.div_(s)
    for fc in linears:
        fc.weight.mul_(s.view(1, -1))


def _attn_qkv(layer: nn.Module) -> List[nn.Linear]:
    return [layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj]


def _mixtral_ffn(layer: nn.Module) -> List[nn.Linear]:
    fcs = [layer.block_sparse_moe.gate]
    for expert in layer.block_sparse_moe.experts:
        fcs += [expert.w1, expert.w3]
    return fcs


# family -> [(norm attribute, linears getter, key of the calibration statistics relative to the layer name)]
_GROUPS: Dict[str, List[Tuple[str, Callable[[nn.Module], Union[nn.Linear, List[nn.Linear]]], str]]] = {
    "transformers": [("self_attn_layer_norm", _attn_qkv, ".self_attn.q_proj"),
                     ("final_layer_norm", lambda m: m.fc1, ".fc1")],
    "llama": [("input_layernorm", _attn_qkv, ".self_attn.q_proj"),
              ("post_attention_layernorm", lambda m: [m.mlp.gate_proj, m.mlp.up_proj], ".mlp.gate_proj")],
    "baichuan": [("input_layernorm", lambda m: m.self_attn.W_pack, ".self_attn.W_pack"),
                 ("post_attention_layernorm", lambda m: [m.mlp.gate_proj, m.mlp.up_proj], ".mlp.gate_proj")],
    "mixtral": [("input_layernorm", _attn_qkv, ".self_attn.q_proj"),
                ("post_attention_layernorm", _mixtral_ffn, ".block_sparse_moe.gate")],
}


def layer_kind(module: nn.Module) -> str:
    """Which decoder-layer family a module is ('' if none), by structure instead of by HF class."""
    if all(hasattr(module, a) for a in ("self_attn_layer_norm", "fc1", "final_layer_norm")):
        return "transformers"  # OPT
    if hasattr(module, "input_layernorm") and hasattr(module, "post_attention_layernorm"):
        if hasattr(getattr(module, "self_attn", None), "W_pack"):
            return "baichuan"
        if hasattr(module, "block_sparse_moe"):
            return "mixtral"
        if hasattr(getattr(module, "mlp", None), "gate_proj"):
            return "llama"
    return ""


@torch.no_grad()
def smooth_lm(model: nn.Module, scales: Dict[str, torch.Tensor], alpha: float = 0.5) -> int:
    """Smooth every (norm, following linears) group of every decoder layer; returns the number of layers touched."""
    touched = 0
    for name, module in model.named_modules():
        kind = layer_kind(module)
        if not kind:
            continue
        for norm_attr, get_linears, stat_key in _GROUPS[kind]:
            smooth_ln_fcs(getattr(module, norm_attr), get_linears(module), scales[name + stat_key], kind, alpha)
        touched += 1
    return touched
