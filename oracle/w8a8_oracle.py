"""CPU oracle for the SmoothQuant W8A8 / FP8 linear hot path.

TEST INFRASTRUCTURE ONLY.  This module is a numpy restatement of the arithmetic of
AniZpZ/AutoSmoothQuant's quantized ``Linear.forward`` and exists to check the CUDA kernels.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it;
the product package (``autosmoothquant_b200``) never does and has no CPU path.

Pinning: the reference ships no tests or golden vectors (its ``tests/__init__.py`` is empty), so
this oracle is pinned against outputs of the reference's own, unmodified Python classes executed
on CPU (``oracle/gen_golden.py`` imports ``/root/reference`` with a stub ``_CUDA`` module that
performs the exact integer matmul and commits the vectors to ``tests/golden/``;
``tests/test_oracle_golden.py`` replays them, and ``tests/test_oracle_vs_reference.py`` fuzzes
against the live reference when it is mounted).

Everything is numpy: float32 arrays carry values that are exactly representable in the
activation dtype ``T`` ('f32' | 'f16' | 'bf16'); rounding to ``T`` and to e4m3 is implemented
here bit by bit, not delegated to torch.

Scalar-division modes (see ``asq_div_mode`` in include/asq.h): torch evaluates
``tensor / python_scalar`` as a true division on CPU and as a multiplication by the fp32
reciprocal on CUDA.  ``div_mode='exact'`` restates the CPU behaviour (what the golden vectors
pin); ``div_mode='reciprocal'`` restates the CUDA behaviour (checked on the GPU box against torch
eager CUDA ops).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

F32 = np.float32
INT8_QMAX = 127.0
E4M3_MAX = 448.0


# ----------------------------------------------------------------------------- dtype rounding
def round_to_bf16(v: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32."""
    shape = np.shape(v)
    v = np.ascontiguousarray(v, dtype=F32).reshape(-1)
    u = v.view(np.uint32).astype(np.uint64)
    bias = 0x7FFF + ((u >> 16) & 1)
    r = ((u + bias) & 0xFFFF0000).astype(np.uint32)
    out = r.view(F32).copy()
    nan = np.isnan(v)
    if nan.any():
        out[nan] = np.nan
    return out.reshape(shape)


def round_to(v: np.ndarray, dtype: str) -> np.ndarray:
    """Round fp32 values to activation dtype T and widen back to fp32."""
    if dtype == "f32":
        return np.asarray(v, dtype=F32)
    if dtype == "f16":
        with np.errstate(over="ignore"):
            return np.asarray(v, dtype=F32).astype(np.float16).astype(F32)
    if dtype == "bf16":
        return round_to_bf16(v)
    raise ValueError(f"unknown dtype {dtype!r}")


def e4m3_encode(v: np.ndarray) -> np.ndarray:
    """fp32 -> float8_e4m3fn bytes, round to nearest even, inputs already clamped to +-448.

    Format: 1 sign, 4 exponent (bias 7), 3 mantissa; no infinities; 0x7F / 0xFF are NaN;
    subnormals have quantum 2^-9; the largest finite value is 448 = 0x7E.
    """
    v = np.asarray(v, dtype=F32)
    sign = (np.signbit(v)).astype(np.uint8) << 7
    a = np.abs(v).astype(np.float64)
    out = np.zeros(v.shape, dtype=np.uint8)
    nan = np.isnan(v)
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.floor(np.log2(np.where(a > 0, a, 1.0))).astype(np.int64)
    e = np.clip(e, -6, 8)  # below 2^-6 everything is subnormal (same quantum as exponent -6)
    quantum = np.exp2((e - 3).astype(np.float64))
    m = np.rint(a / quantum)  # in [0, 16]; 16 means carry into the next binade
    val = m * quantum
    val = np.minimum(val, E4M3_MAX)
    # re-derive fields from the rounded value
    with np.errstate(divide="ignore", invalid="ignore"):
        e2 = np.floor(np.log2(np.where(val > 0, val, 1.0))).astype(np.int64)
    is_sub = val < 2.0 ** -6
    e2 = np.where(is_sub, -6, e2)
    mant = np.where(is_sub, np.rint(val / 2.0 ** -9), np.rint(val / np.exp2((e2 - 3).astype(np.float64))) - 8)
    exp_field = np.where(is_sub, 0, e2 + 7)
    out = (exp_field.astype(np.uint8) << 3) | mant.astype(np.uint8)
    out = out | sign
    out = np.where(nan, np.uint8(0x7F), out).astype(np.uint8)
    return out


def e4m3_decode(b: np.ndarray) -> np.ndarray:
    """float8_e4m3fn bytes -> fp32 (exact)."""
    b = np.asarray(b, dtype=np.uint8)
    sign = np.where(b & 0x80, -1.0, 1.0)
    ex = ((b >> 3) & 0xF).astype(np.int64)
    mant = (b & 0x7).astype(np.float64)
    val = np.where(ex == 0, mant * 2.0 ** -9, (8 + mant) * np.exp2((ex - 10).astype(np.float64)))
    val = np.where((b & 0x7F) == 0x7F, np.nan, val)
    return (sign * val).astype(F32)


# ----------------------------------------------------------------------------- helpers
def _scalar_div(v: np.ndarray, s: float, dtype: str, div_mode: str) -> np.ndarray:
    """``tensor_T / python_scalar`` as torch computes it: fp32 math, one rounding to T."""
    v = np.asarray(v, dtype=F32)
    s32 = F32(s)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if div_mode == "exact":
            r = v / s32
        elif div_mode == "reciprocal":
            r = v * (F32(1.0) / s32)
        else:
            raise ValueError(f"unknown div_mode {div_mode!r}")
    return round_to(r, dtype)


def sat_i8(v: np.ndarray) -> np.ndarray:
    """``.clamp(-128, 127).to(torch.int8)``; NaN -> 0 (what torch yields on CPU and CUDA)."""
    v = np.asarray(v, dtype=F32)
    v = np.where(np.isnan(v), F32(0), v)
    return np.clip(v, -128, 127).astype(np.int8)


def int8_gemm_i32(a: np.ndarray, w: np.ndarray) -> np.ndarray:
    """C[M,N] = A[M,K] . W[N,K]^T, exact int32.

    Follows csrc/int8gemm/bindings.cpp:69-84 (m = input.size(0), n = weight.size(0),
    k = input.size(1)) and cublasINT8MMWrapper.cc:224-354 (CUBLAS_COMPUTE_32I, alpha=1, beta=0).
    float64 BLAS is exact here: |acc| <= K * 128 * 128 < 2^53.
    """
    a = np.asarray(a)
    w = np.asarray(w)
    assert a.dtype == np.int8 and w.dtype == np.int8 and a.shape[1] == w.shape[1]
    acc = a.astype(np.float64) @ w.astype(np.float64).T
    return acc.astype(np.int64).astype(np.int32)


# ----------------------------------------------------------------------------- offline weight quant
def quantize_per_tensor_absmax(w: np.ndarray, dtype: str = "f32") -> Tuple[np.ndarray, np.float32]:
    """layers/functional/quantization.py:9-18.

    scale = max|W| / 127 in W's dtype; on CPU the weight is widened to fp32 first (``t.float()``),
    then ``div_(scale).round_()`` and ``.to(int8)`` without a clamp.
    """
    w = np.asarray(w, dtype=F32)
    scale = F32(round_to(np.abs(w).max() / F32(127), dtype).reshape(()))  # 0-dim tensor in W's dtype
    q = np.rint(w / scale)
    return q.astype(np.int8), scale


# ----------------------------------------------------------------------------- activation quant
def quantize_act_int8(
    x: np.ndarray, dtype: str, act_mode: str, quant_scale: float = 1.0, div_mode: str = "exact"
) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """The three INT8 prologues of layers/nn/linear.py.

    'round'      x.round().clamp(-128,127).to(int8)                                     :95
    'scale'      (x / quant_scale.item()).round().clamp().to(int8), division in T       :290-292
    'per-token'  s = x.abs().max(-1, keepdim)[0].div(127.0).to(fp32)  (division in T)   :88-90 / :283-285
                 q = (x / s).round().clamp().to(int8)  (x promoted to fp32)             :92 / :292
    Returns (q int8 [M,K], row_scale fp32 [M] or None).
    """
    x = np.asarray(x, dtype=F32)
    if act_mode == "round":
        return sat_i8(np.rint(x)), None
    if act_mode == "scale":
        return sat_i8(np.rint(_scalar_div(x, quant_scale, dtype, div_mode))), None
    if act_mode == "per-token":
        amax = np.abs(x).max(axis=-1, keepdims=True) if x.shape[0] else np.zeros((0, 1), F32)
        s = _scalar_div(amax, INT8_QMAX, dtype, div_mode)
        with np.errstate(divide="ignore", invalid="ignore"):
            v = x / s
        return sat_i8(np.rint(v)), s.reshape(-1).astype(F32)
    raise ValueError(f"unknown act_mode {act_mode!r}")


def quantize_act_fp8(
    x: np.ndarray, dtype: str, act_mode: str, in_scale: float = 1.0, div_mode: str = "exact"
) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """FP8 prologues of layers/functional/quantization.py.

    'per-token'  per_token_quantize_fp8  :173-191   s = absmax.div(448).to(fp32); q = e4m3(clamp(x / s))
    'scale'      static_per_tensor_quantize_fp8 :208-211   q = e4m3(clamp(x / inv_scale)), division in T
    'per-tensor' per_tensor_quantize_fp8 :144-170   s = max(|min|,|max|) / 448 in T (0-dim), q = e4m3(clamp(x / s))
    Returns (e4m3 bytes [M,K], scale: fp32 [M] | 0-dim fp32 | None).
    """
    x = np.asarray(x, dtype=F32)
    if act_mode == "per-token":
        amax = np.abs(x).max(axis=-1, keepdims=True)
        s = _scalar_div(amax, E4M3_MAX, dtype, div_mode)
        with np.errstate(divide="ignore", invalid="ignore"):
            v = x / s
        return e4m3_encode(np.clip(v, -E4M3_MAX, E4M3_MAX)), s.reshape(-1).astype(F32)
    if act_mode == "scale":
        v = _scalar_div(x, in_scale, dtype, div_mode)
        return e4m3_encode(np.clip(v, -E4M3_MAX, E4M3_MAX)), None
    if act_mode == "per-tensor":
        amax = np.maximum(np.abs(x.min()), np.abs(x.max())) if x.size else F32(16.0)
        s = _scalar_div(np.asarray(amax, F32), E4M3_MAX, dtype, div_mode)  # tensor / python float
        with np.errstate(divide="ignore", invalid="ignore"):
            v = round_to(x / F32(s), dtype)  # T tensor / 0-dim T tensor: true division, rounded to T
        return e4m3_encode(np.clip(v, -E4M3_MAX, E4M3_MAX)), np.asarray(s, F32)
    raise ValueError(f"unknown act_mode {act_mode!r}")


# ----------------------------------------------------------------------------- INT8 forwards
def _dequant(acc: np.ndarray, factor, bias: Optional[np.ndarray], dtype: str) -> np.ndarray:
    """``out = dequant_scale * out (+ bias)``; ``.to(dtype)``  — linear.py:104-105.

    int32 -> fp32 conversion is round-to-nearest-even; the multiply and the add are separate fp32
    roundings (eager torch launches, no FMA)."""
    y = np.asarray(factor, dtype=F32) * acc.astype(F32)
    if bias is not None:
        y = y + np.asarray(bias, dtype=F32)
    return round_to(y.astype(F32), dtype)


def w8a8_linear(
    x: np.ndarray,
    dtype: str,
    weight: np.ndarray,
    dequant_scale: float,
    act_quant: str = "per-tensor",
    bias: Optional[np.ndarray] = None,
    quant_scale: Optional[float] = None,
    div_mode: str = "exact",
) -> np.ndarray:
    """W8A8BFP32OFP32Linear.forward (linear.py:83-106) when ``quant_scale is None`` and
    W8A8BFP32OFP32LinearWithQuantScale.forward (linear.py:278-302) when it is given
    (per-tensor only; per-token ignores it exactly like the reference).

    x: [..., K] fp32 array holding T-representable values; returns [..., N] fp32 holding T values.
    """
    x = np.asarray(x, dtype=F32)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if act_quant == "per-token":
        q, s = quantize_act_int8(x2, dtype, "per-token", div_mode=div_mode)
        factor = (F32(dequant_scale) * s).reshape(-1, 1)  # python float * fp32 tensor
    elif act_quant == "per-tensor":
        mode = "round" if quant_scale is None else "scale"
        q, _ = quantize_act_int8(x2, dtype, mode, 1.0 if quant_scale is None else quant_scale, div_mode)
        factor = F32(dequant_scale)
    else:
        raise AssertionError('"act_quant must be "per-token" or "per-tensor"')
    acc = int8_gemm_i32(q, weight)
    return _dequant(acc, factor, bias, dtype).reshape(*lead, weight.shape[0])


def w8a8_qkv_linear(
    x: np.ndarray,
    dtype: str,
    weight: np.ndarray,
    qkv_size: Sequence[int],
    q_scale: float,
    k_scale: float,
    v_scale: float,
    act_quant: str = "per-tensor",
    bias: Optional[np.ndarray] = None,
    div_mode: str = "exact",
) -> np.ndarray:
    """W8A8BFP32OFP32QKVLinear.forward (linear.py:172-208): one GEMM, three column blocks each with
    its own scalar dequant scale (times the row scale for per-token) and bias slice."""
    x = np.asarray(x, dtype=F32)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if act_quant == "per-token":
        q, s = quantize_act_int8(x2, dtype, "per-token", div_mode=div_mode)
        s = s.reshape(-1, 1)
    else:
        q, _ = quantize_act_int8(x2, dtype, "round")
        s = None
    acc = int8_gemm_i32(q, weight).astype(F32)
    outs = []
    col = 0
    for width, sc in zip(qkv_size, (q_scale, k_scale, v_scale)):
        f = F32(sc) * s if s is not None else F32(sc)
        blk = (f * acc[:, col:col + width]).astype(F32)
        if bias is not None:
            blk = blk + np.asarray(bias, F32)[col:col + width]
        outs.append(blk.astype(F32))
        col += width
    y = np.concatenate(outs, axis=-1)
    return round_to(y, dtype).reshape(*lead, weight.shape[0])


# ----------------------------------------------------------------------------- FP8 forwards
def fp8_linear_exact(
    q_bytes: np.ndarray, w_bytes: np.ndarray, a_scale, w_scale: float, bias: Optional[np.ndarray] = None
) -> np.ndarray:
    """fp64 evaluation of sum_k q*w * sA * sW (+ bias): the value both the reference's
    dequantise-then-F.linear (linear.py:363-368) and the tensor-core kernel approximate."""
    a = e4m3_decode(q_bytes).astype(np.float64)
    w = e4m3_decode(w_bytes).astype(np.float64)
    y = (a @ w.T) * np.asarray(a_scale, np.float64).reshape(-1, 1) * float(w_scale)
    if bias is not None:
        y = y + np.asarray(bias, np.float64)
    return y


def fp8_linear_reference_math(
    x: np.ndarray, dtype: str, w_bytes: np.ndarray, w_scale: float, act_quant: str = "per-token",
    in_scale: float = 1.0, bias: Optional[np.ndarray] = None, div_mode: str = "exact", out_scale: float = 0.0,
) -> np.ndarray:
    """FP8LinearDynamic / FP8LinearStatic forward as the reference computes it (easy_fp8_gemm,
    linear.py:336-369): ``F.linear(A.to(dt) * sA, W.to(dt) * sW, bias)`` in the activation dtype, optionally
    followed by the output fake-quantisation of FP8LinearStatic (linear.py:562-564).  The reference raises a
    dtype mismatch for per-token fp16/bf16 inputs, so per-token is fp32-only; the
    accumulation order of a BLAS GEMM is not pinned, so compare with a tolerance."""
    x = np.asarray(x, dtype=F32)
    mode = {"per-token": "per-token", "static": "scale", "per-tensor": "per-tensor"}[act_quant]
    q, s = quantize_act_fp8(x.reshape(-1, x.shape[-1]), dtype, mode, in_scale, div_mode)
    a_scale = F32(in_scale) if mode == "scale" else s
    a = round_to(e4m3_decode(q) * np.asarray(a_scale, F32).reshape(-1, 1), dtype)
    w = round_to(e4m3_decode(w_bytes) * F32(w_scale), dtype)
    y = a.astype(np.float64) @ w.astype(np.float64).T
    if bias is not None:
        y = y + np.asarray(bias, np.float64)
    y = round_to(y.astype(F32), dtype)
    if out_scale:
        t = _scalar_div(y, out_scale, dtype, div_mode)
        y = round_to(e4m3_decode(e4m3_encode(np.clip(t, -E4M3_MAX, E4M3_MAX))) * F32(out_scale), dtype)
    return y.reshape(*x.shape[:-1], w_bytes.shape[0])


# ----------------------------------------------------------------------------- tensor-parallel restatement
def tp_row_parallel_int32(q: np.ndarray, weight: np.ndarray, world: int) -> np.ndarray:
    """Row-parallel INT8 GEMM: rank r holds W[:, rK/p:(r+1)K/p] and the matching slice of q; the int32
    partial products are summed (the all-reduce).  Must equal int8_gemm_i32(q, weight) exactly."""
    K = q.shape[1]
    assert K % world == 0
    kp = K // world
    parts = [int8_gemm_i32(np.ascontiguousarray(q[:, r * kp:(r + 1) * kp]),
                           np.ascontiguousarray(weight[:, r * kp:(r + 1) * kp])) for r in range(world)]
    return np.sum(np.stack(parts).astype(np.int64), axis=0).astype(np.int32)


# ----------------------------------------------------------------------------- producer-side fusions (asq_glue.cu)
def add_rmsnorm_quant(x: np.ndarray, delta: Optional[np.ndarray], weight: np.ndarray, eps: float, dtype: str):
    """Residual add + HF LlamaRMSNorm.forward (inherited by the reference's QuantizedLlamaRMSNorm,
    models/llama.py:27-37, weight already divided by the input scale) + the rounding the per-tensor Linear
    applies to its input (linear.py:95).  Returns (x + delta in T, h in T, int8 sat(rint(h)))."""
    x = np.asarray(x, F32)
    if delta is not None:
        x = round_to(x + np.asarray(delta, F32), dtype)
    var = np.mean(x.astype(np.float64) ** 2, axis=-1, keepdims=True).astype(F32)
    n = round_to(x * (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32), dtype)
    h = round_to(np.asarray(weight, F32) * n, dtype)
    return x, h, sat_i8(np.rint(h))


def silu_mul_quant(gate: np.ndarray, up: np.ndarray, quant_scale: float, dtype: str, div_mode: str = "exact"):
    """a = T(T(silu(gate)) * up) (HF LlamaMLP: act_fn(gate_proj(x)) * up_proj(x)) and the int8 tensor
    W8A8BFP32OFP32LinearWithQuantScale derives from it (linear.py:290-292)."""
    g = np.asarray(gate, F32)
    s = round_to((g.astype(np.float64) / (1.0 + np.exp(-g.astype(np.float64)))).astype(F32), dtype)
    a = round_to(s * np.asarray(up, F32), dtype)
    q = sat_i8(np.rint(_scalar_div(a, quant_scale, dtype, div_mode)))
    return a, q


def rope_rotate_half(x: np.ndarray, cos: np.ndarray, sin: np.ndarray, dtype: str) -> np.ndarray:
    """HF apply_rotary_pos_emb for one tensor: T(T(x*cos) + T(rotate_half(x)*sin)); x [..., S, hd]."""
    x = np.asarray(x, F32)
    half = x.shape[-1] // 2
    rot = np.concatenate([-x[..., half:], x[..., :half]], axis=-1)
    return round_to(round_to(x * cos, dtype) + round_to(rot * sin, dtype), dtype)
