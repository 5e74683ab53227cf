"""Offline quantisation helpers (weights, calibration-time tensors).

These run once when a float model is converted (``from_float``); they are plain torch and are not
on the inference hot path — at run time the activation quantisation they describe is executed by
the fused kernel's prologue.  Semantics follow the reference's
``autosmoothquant/layers/functional/quantization.py`` (line numbers below), written without the
in-place mutation of the source tensor the reference performs.
"""
from __future__ import annotations

from typing import Tuple

import torch

E4M3_MAX = torch.finfo(torch.float8_e4m3fn).max  # 448


@torch.no_grad()
def quantize_per_tensor_absmax(t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(int8 weight, 0-dim scale in t's dtype); scale = max|t| / 127, no clamp (quantization.py:9-18)."""
    scale = t.abs().max() / 127
    q = (t.float() if not t.is_cuda else t).div(scale).round()
    return q.to(torch.int8), scale


@torch.no_grad()
def quantize_weight_per_channel_absmax(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-output-channel int8 weights, scales [N,1] (quantization.py:38-49)."""
    scales = (w.abs().max(dim=1)[0] / 127).view(-1, 1)
    q = (w.float() if not w.is_cuda else w).div(scales).round().clamp(-128, 127)
    return q.to(torch.int8), scales


@torch.no_grad()
def per_tensor_quantize_fp8(tensor: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(e4m3 tensor, 0-dim scale) with scale = max(|min|, |max|) / 448 (quantization.py:144-170);
    an empty tensor (empty MoE expert) gets the reference's placeholder range of +-16."""
    if tensor.numel() == 0:
        amax = torch.tensor(16.0, dtype=tensor.dtype)
    else:
        lo, hi = tensor.aminmax()
        amax = torch.maximum(lo.abs(), hi.abs())
    scale = amax / E4M3_MAX
    q = (tensor / scale).clamp(min=-E4M3_MAX, max=E4M3_MAX).to(torch.float8_e4m3fn)
    return q, scale


@torch.no_grad()
def per_token_quantize_fp8(tensor: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(e4m3 tensor, fp32 scales [...,1]) with one scale per row (quantization.py:173-191)."""
    assert tensor.numel() > 0
    scale = tensor.abs().max(dim=-1, keepdim=True)[0].div(E4M3_MAX).to(torch.float32)
    q = (tensor / scale).clamp(min=-E4M3_MAX, max=E4M3_MAX).to(torch.float8_e4m3fn)
    return q, scale


@torch.no_grad()
def static_per_tensor_quantize_fp8(tensor: torch.Tensor, inv_scale) -> torch.Tensor:
    """e4m3(clamp(tensor / inv_scale)) (quantization.py:208-211)."""
    return (tensor / inv_scale).clamp(min=-E4M3_MAX, max=E4M3_MAX).to(torch.float8_e4m3fn)


def dtype_byte_size(dtype: torch.dtype) -> float:
    """Bytes per element of `dtype`, float8 types included.  The reference monkeypatches
    ``transformers.modeling_utils.dtype_byte_size`` with a regex-based version (quantization.py:126-136) because the
    transformers release it pins mis-sizes ``torch.float8_e4m3fn`` when ``save_pretrained`` shards an FP8 checkpoint;
    this one asks torch (``bool`` counts as one bit, as in transformers)."""
    if dtype == torch.bool:
        return 1 / 8
    if not isinstance(dtype, torch.dtype):
        raise ValueError(f"`dtype` is not a valid dtype: {dtype}.")
    return torch.empty((), dtype=dtype).element_size()


def install_dtype_byte_size_patch() -> bool:
    """Same side effect as importing the reference's quantization module: transformers' checkpoint sharding sizes FP8
    tensors correctly.  Returns False when transformers is not importable (nothing to patch)."""
    try:
        import transformers.modeling_utils as mu
    except Exception:  # noqa: BLE001
        return False
    mu.dtype_byte_size = dtype_byte_size
    return True


install_dtype_byte_size_patch()
