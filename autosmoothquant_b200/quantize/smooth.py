"""Activation smoothing (SmoothQuant): migrate quantisation difficulty from activations to weights.

Mirrors ``autosmoothquant/quantize/smooth.py:10-93`` of the reference: ``smooth_ln_fcs`` is the same arithmetic
(per input channel ``s = act_absmax**alpha / weight_absmax**(1-alpha)``, norm weight (and bias) divided by s, the
following Linear weights multiplied by s); ``smooth_lm`` walks the decoder layers.  The reference dispatches on
HF classes of transformers 4.42 (``OPTDecoderLayer`` ...); here layers are recognised by their attribute
structure, so the same code serves the installed transformers and the bench's own stack.  Offline, not timed.
"""
from __future__ import annotations

from typing import Dict, List, Union

import torch
from torch import nn


@torch.no_grad()
def smooth_ln_fcs(ln: nn.Module, fcs: Union[nn.Linear, List[nn.Linear]], act_scales: torch.Tensor,
                  model_type: str = "transformers", alpha: float = 0.5) -> None:
    """smooth.py:10-40.  ``model_type == "transformers"`` means a LayerNorm with bias (OPT); every other type
    (llama / mixtral / baichuan) is an RMSNorm with a weight only."""
    if not isinstance(fcs, list):
        fcs = [fcs]
    for fc in fcs:
        assert isinstance(fc, nn.Linear)
        assert ln.weight.numel() == fc.in_features == act_scales.numel()
    if model_type == "transformers":
        assert isinstance(ln, nn.LayerNorm)
    device, dtype = fcs[0].weight.device, fcs[0].weight.dtype
    act_scales = act_scales.to(device=device, dtype=dtype)
    weight_scales = torch.cat([fc.weight.abs().max(dim=0, keepdim=True)[0] for fc in fcs], dim=0)
    weight_scales = weight_scales.max(dim=0)[0].clamp(min=1e-5)
    scales = (act_scales.pow(alpha) / weight_scales.pow(1 - alpha)).clamp(min=1e-5).to(device).to(dtype)
    ln.weight.div_(scales)
    if model_type == "transformers":
        ln.bias.div_(scales)
    for fc in fcs:
        fc.weight.mul_(scales.view(1, -1))


def layer_kind(module: nn.Module) -> str:
    """Which decoder-layer family a module is ('' if none), by structure instead of by HF class."""
    if hasattr(module, "self_attn_layer_norm") and hasattr(module, "fc1") and hasattr(module, "final_layer_norm"):
        return "transformers"  # OPT
    if hasattr(module, "input_layernorm") and hasattr(module, "post_attention_layernorm"):
        attn = getattr(module, "self_attn", None)
        if attn is not None and hasattr(attn, "W_pack"):
            return "baichuan"
        if hasattr(module, "block_sparse_moe"):
            return "mixtral"
        if hasattr(module, "mlp") and hasattr(module.mlp, "gate_proj"):
            return "llama"
    return ""


@torch.no_grad()
def smooth_lm(model: nn.Module, scales: Dict[str, torch.Tensor], alpha: float = 0.5) -> int:
    """smooth.py:42-93: smooth (norm, following linears) pairs of every decoder layer; returns the layer count."""
    n = 0
    for name, module in model.named_modules():
        kind = layer_kind(module)
        if kind == "transformers":
            qkv = [module.self_attn.q_proj, module.self_attn.k_proj, module.self_attn.v_proj]
            smooth_ln_fcs(module.self_attn_layer_norm, qkv, scales[name + ".self_attn.q_proj"], "transformers", alpha)
            smooth_ln_fcs(module.final_layer_norm, module.fc1, scales[name + ".fc1"], "transformers", alpha)
        elif kind == "llama":
            qkv = [module.self_attn.q_proj, module.self_attn.k_proj, module.self_attn.v_proj]
            smooth_ln_fcs(module.input_layernorm, qkv, scales[name + ".self_attn.q_proj"], "llama", alpha)
            smooth_ln_fcs(module.post_attention_layernorm, [module.mlp.gate_proj, module.mlp.up_proj],
                          scales[name + ".mlp.gate_proj"], "llama", alpha)
        elif kind == "baichuan":
            smooth_ln_fcs(module.input_layernorm, module.self_attn.W_pack, scales[name + ".self_attn.W_pack"], "baichuan", alpha)
            smooth_ln_fcs(module.post_attention_layernorm, [module.mlp.gate_proj, module.mlp.up_proj],
                          scales[name + ".mlp.gate_proj"], "baichuan", alpha)
        elif kind == "mixtral":
            qkv = [module.self_attn.q_proj, module.self_attn.k_proj, module.self_attn.v_proj]
            smooth_ln_fcs(module.input_layernorm, qkv, scales[name + ".self_attn.q_proj"], "mixtral", alpha)
            fcs = [module.block_sparse_moe.gate]
            for expert in module.block_sparse_moe.experts:
                fcs += [expert.w1, expert.w3]
            smooth_ln_fcs(module.post_attention_layernorm, fcs, scales[name + ".block_sparse_moe.gate"], "mixtral", alpha)
        else:
            continue
        n += 1
    return n
