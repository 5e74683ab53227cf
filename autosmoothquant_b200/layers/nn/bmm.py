"""INT8 batched-matmul modules, interface-compatible with the reference's ``layers/nn/bmm.py`` (three classes,
``forward(a, b)`` with a [B, M, K] and b [B, N, K] int8, ``from_scale`` constructors, an ``a`` buffer holding alpha).

All three share one implementation: ``asq_i8bmm`` (include/asq.h) computes the batch on the grouped tcgen05 path and
only the epilogue differs — raw int32, alpha-scaled float32, or alpha-scaled, rounded and saturated int8
(csrc/kernels/bmm.cu:10-211 in the reference).
"""
from __future__ import annotations

import torch
from torch import nn

from ... import _lib


class _Int8BatchedMatmul(nn.Module):
    """c[b] = epilogue(alpha * a[b] @ b[b]^T); subclasses fix the output type."""

    OUT_DTYPE: torch.dtype = torch.int32
    HAS_ALPHA = True

    def __init__(self, alpha: float = 1.0):
        super().__init__()
        if self.HAS_ALPHA:
            self.register_buffer("a", torch.as_tensor(alpha))

    def _alpha(self) -> float:
        return float(self.a) if self.HAS_ALPHA else 1.0

    def _set_alpha(self, value) -> "_Int8BatchedMatmul":
        self.a = value if torch.is_tensor(value) else torch.as_tensor(value)
        return self

    @torch.no_grad()
    def forward(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return _lib.i8bmm(a, b, self.OUT_DTYPE, self._alpha())


class BMM_S8T_S8N_S8T(_Int8BatchedMatmul):
    """int8 out: sat_i8(rint(alpha * acc)), alpha = a_scale * b_scale / output_scale."""

    OUT_DTYPE = torch.int8

    @classmethod
    def from_scale(cls, a_scale, b_scale, output_scale):
        return cls(1.0)._set_alpha(a_scale * b_scale / output_scale)


class BMM_S8T_S8N_F32T(_Int8BatchedMatmul):
    """float32 out: alpha * acc, alpha = a_scale * b_scale."""

    OUT_DTYPE = torch.float32

    @classmethod
    def from_scale(cls, a_scale, b_scale):
        return cls(1.0)._set_alpha(a_scale * b_scale)


class BMM_S8T_S8N_S32T(_Int8BatchedMatmul):
    """int32 out: the exact accumulators; no scale."""

    OUT_DTYPE = torch.int32
    HAS_ALPHA = False

    def __init__(self):
        super().__init__()
