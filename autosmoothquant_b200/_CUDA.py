"""Source-compatible stand-in for the reference's pybind11 module ``autosmoothquant._CUDA``.

The reference exposes one class, ``I8CUGEMM`` (csrc/int8gemm/bindings.cpp:145-155), whose five
methods wrap cuBLASLt INT8 GEMMs.  Here the same five methods call the hand-written sm_100a
kernels through the C ABI (include/asq.h).  Differences from the reference, all deliberate:

* stateless — no process-wide mutex (cublasINT8MMWrapper.cc:228) and no stream captured at
  construction (bindings.cpp:13): every call uses the current stream of the tensors' device;
* arguments are validated (dtype, shape, contiguity, device) and errors raise ``RuntimeError``
  instead of being dropped (cublasINT8MMWrapper.cc:343-346 ignores cuBLASLt's status);
* there is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _lib


class I8CUGEMM:
    """INT8 GEMM entry points: ``input`` [M,K] int8, ``weight`` [N,K] int8 (K contiguous)."""

    def linear_a8_w8_o32_(self, input: torch.Tensor, weight: torch.Tensor, out: torch.Tensor) -> None:
        """out[M,N] (int32) = input @ weight^T, in place (bindings.cpp:69-84) — the hot-path call."""
        _lib.i8gemm_o32(input, weight, out)

    def linear_a8_w8_o32(self, input: torch.Tensor, weight: torch.Tensor, out: torch.Tensor) -> None:
        """Same contract (bindings.cpp:52-67).  The reference's variant expects COL32-interleaved
        operands nobody produces; row-major operands are the only layout its callers ever pass."""
        _lib.i8gemm_o32(input, weight, out)

    def linear_a8_w8_o8(self, input: torch.Tensor, weight: torch.Tensor, out: torch.Tensor, alpha: float) -> None:
        """out (int8) = sat(rint(alpha * (input @ weight^T)))  (bindings.cpp:86-102)."""
        _lib.i8gemm_epi(input, weight, out, alpha, 0.0)

    def linear_a8_w8_o8_(self, input: torch.Tensor, weight: torch.Tensor, out: torch.Tensor, alpha: float,
                         beta: float = 0.0) -> None:
        """out (int8) = sat(rint(alpha * (input @ weight^T) + beta * out_in))  (bindings.cpp:104-121).
        With beta == 0 (the default and the only value the reference's callers use) ``out`` is write-only and the call
        is one fused launch.  beta != 0 reads the previous contents of ``out`` as the cuBLASLt C operand
        (cublasINT8MMWrapper.cc:537-672): the GEMM launch emits fp32 ``alpha * acc`` and one elementwise pass applies
        ``+ beta * out_in``, rint and saturation — the same fp32 operations in the same order as the fused alpha / beta
        epilogue of ``linear_a8_w8_b8_o8_``, on an entry point nothing in the reference calls."""
        if beta == 0.0:
            _lib.i8gemm_epi(input, weight, out, alpha, 0.0)
            return
        if out.dtype != torch.int8 or tuple(out.shape) != (input.shape[0], weight.shape[0]):
            raise ValueError("linear_a8_w8_o8_: out must be int8 [M,N]")
        scaled = torch.empty(out.shape, dtype=torch.float32, device=out.device)
        _lib.i8gemm_epi(input, weight, scaled, alpha, 0.0)
        out.copy_(torch.clamp(torch.round(scaled + torch.tensor(beta, dtype=torch.float32, device=out.device) * out.to(torch.float32)),
                              -128, 127).to(torch.int8))

    def linear_a8_w8_b8_o8_(self, input: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, alpha: float,
                            beta: float) -> torch.Tensor:
        """returns int8 [M,N] = sat(rint(alpha * (input @ weight^T) + beta * bias[n]))  (bindings.cpp:123-142;
        the reference materialises ``bias.repeat(M, 1)`` as the C operand, here the epilogue reads bias[n])."""
        out = torch.empty((input.shape[0], weight.shape[0]), dtype=torch.int8, device=input.device)
        _lib.i8gemm_epi(input, weight, out, alpha, beta, bias=bias.reshape(-1).contiguous())
        return out


# ---- csrc/kernels/bmm.cu (exported by the reference's legacy `_CUDA` build, used by layers/nn/bmm.py)
def bmm_s8t_s8n_s8t(a: torch.Tensor, b: torch.Tensor, alpha: float) -> torch.Tensor:
    """int8 [B,M,N] = sat(rint(alpha * a[B,M,K] @ b[B,N,K]^T))  (bmm.cu:82-148, LinearCombinationClamp)."""
    return _lib.i8bmm(a, b, torch.int8, alpha)


def bmm_s8t_s8n_f32t(a: torch.Tensor, b: torch.Tensor, alpha: float) -> torch.Tensor:
    """float32 [B,M,N] = alpha * a @ b^T  (bmm.cu:10-80)."""
    return _lib.i8bmm(a, b, torch.float32, alpha)


def bmm_s8t_s8n_s32t(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """int32 [B,M,N] = a @ b^T, exact  (bmm.cu:150-211)."""
    return _lib.i8bmm(a, b, torch.int32, 1.0)
