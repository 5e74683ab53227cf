"""ctypes binding of ``libasq_b200.so`` (the C ABI declared in ``include/asq.h``).

PyTorch is used only for device memory and streams: every compute call hands raw device
pointers, sizes and the current CUDA stream to the library.  There is no CPU or eager fallback:
if the library is missing, fails to load, or the tensors are not on a CUDA device, the call
raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from pathlib import Path
from typing import Optional, Tuple

import torch

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / os.environ.get("ASQ_LIB_NAME", "libasq_b200.so")  # ASQ_LIB_NAME: experiment builds only

# enums of include/asq.h
ASQ_F32, ASQ_F16, ASQ_BF16, ASQ_I32, ASQ_I8 = 0, 1, 2, 3, 4
ACT_ROUND, ACT_SCALE, ACT_PER_TOKEN, ACT_PER_TENSOR_DYNAMIC, ACT_ROW_SCALE_GIVEN = 0, 1, 2, 3, 4
DIV_RECIPROCAL, DIV_EXACT = 0, 1
EPI_RELU = 1

_DTYPE_CODE = {
    torch.float32: ASQ_F32,
    torch.float16: ASQ_F16,
    torch.bfloat16: ASQ_BF16,
    torch.int32: ASQ_I32,
    torch.int8: ASQ_I8,
}

# every symbol include/asq.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = (
    "asq_version",
    "asq_last_error",
    "asq_device_supported",
    "asq_workspace_bytes",
    "asq_w8a8_linear",
    "asq_fp8_linear",
    "asq_fp8_linear_cs",
    "asq_i8gemm_o32",
    "asq_i8gemm_epi",
    "asq_quantize_act",
    "asq_w8a8_linear_q8",
    "asq_w8a8_gateup_swiglu_q8",
    "asq_w8a8_linear_q8_rope",
    "asq_w8a8_linear_res",
    "asq_w8a8_rmsnorm_linear_rope",
    "asq_w8a8_rmsnorm_gateup_swiglu",
    "asq_w8a8_linear_q8_res",
    "asq_w8a8_grouped_linear",
    "asq_i8bmm",
    "asq_ar_buffer_bytes",
    "asq_w8a8_linear_q8_allreduce",
    "asq_q8_linear_allreduce_nvls",
    "asq_nvls_probe",
    "asq_dev_alloc",
    "asq_dev_free",
    "asq_ipc_export",
    "asq_ipc_open",
    "asq_ipc_close",
    "asq_add_rmsnorm_quant",
    "asq_silu_mul_quant",
    "asq_rope_inplace",
)

_lib = None
_lib_lock = threading.Lock()
_launches = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)

# How `tensor / python_scalar` is evaluated (see asq_div_mode in include/asq.h).  The default
# reproduces the reference on its only supported device (CUDA); tests switch to DIV_EXACT to
# compare with the CPU oracle that is pinned against the reference's Python executed on CPU.
_default_div_mode = DIV_EXACT if os.environ.get("ASQ_DIV_MODE", "").lower() == "exact" else DIV_RECIPROCAL


def set_div_mode(mode: int) -> int:
    """Set the process-wide scalar-division mode; returns the previous one."""
    global _default_div_mode
    if mode not in (DIV_RECIPROCAL, DIV_EXACT):
        raise ValueError("div mode must be DIV_RECIPROCAL or DIV_EXACT")
    prev, _default_div_mode = _default_div_mode, mode
    return prev


def get_div_mode() -> int:
    return _default_div_mode


def launch_count() -> int:
    return _launches


def load():
    """Load the shared library (once) and declare the C signatures."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m autosmoothquant_b200.build` "
                "(there is no CPU / eager fallback for the W8A8 path)"
            )
        lib = ctypes.CDLL(str(LIB_PATH))
        c_vp, c_i, c_i64, c_f, c_sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
        lib.asq_version.restype = c_i
        lib.asq_version.argtypes = []
        lib.asq_last_error.restype = ctypes.c_char_p
        lib.asq_last_error.argtypes = []
        lib.asq_device_supported.restype = c_i
        lib.asq_device_supported.argtypes = []
        lib.asq_workspace_bytes.restype = c_sz
        lib.asq_workspace_bytes.argtypes = [c_i64, c_i64]
        lib.asq_w8a8_linear.restype = c_i
        lib.asq_w8a8_linear.argtypes = [c_vp, c_i, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_i, c_f, c_f,
                                        c_vp, c_vp, c_i, c_vp, c_sz, c_vp]
        lib.asq_fp8_linear.restype = c_i
        lib.asq_fp8_linear.argtypes = [c_vp, c_i, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_i, c_f, c_f, c_f,
                                       c_vp, c_i, c_vp, c_sz, c_vp]
        lib.asq_fp8_linear_cs.restype = c_i
        lib.asq_fp8_linear_cs.argtypes = [c_vp, c_i, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_i, c_vp,
                                          c_vp, c_i, c_vp, c_sz, c_vp]
        lib.asq_i8gemm_o32.restype = c_i
        lib.asq_i8gemm_o32.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_sz, c_vp]
        lib.asq_i8gemm_epi.restype = c_i
        lib.asq_i8gemm_epi.argtypes = [c_vp, c_vp, c_vp, c_i, c_vp, c_i, c_i64, c_i64, c_i64, c_f, c_f, c_i, c_vp, c_sz, c_vp]
        lib.asq_quantize_act.restype = c_i
        lib.asq_quantize_act.argtypes = [c_vp, c_i, c_vp, c_vp, c_i64, c_i64, c_i, c_f, c_i, c_i, c_vp]
        lib.asq_w8a8_linear_q8.restype = c_i
        lib.asq_w8a8_linear_q8.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_f, c_vp, c_vp, c_sz, c_vp]
        lib.asq_w8a8_gateup_swiglu_q8.restype = c_i
        lib.asq_w8a8_gateup_swiglu_q8.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_i64, c_i64, c_i64, c_f, c_f,
                                                  c_vp, c_f, c_i, c_vp]
        lib.asq_w8a8_rmsnorm_linear_rope.restype = c_i
        lib.asq_w8a8_rmsnorm_linear_rope.argtypes = [c_vp, c_i, c_vp, c_f, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f, c_vp,
                                                     c_vp, c_vp, c_i64, c_i64, c_i64, c_i, c_vp, c_sz, c_vp]
        lib.asq_w8a8_rmsnorm_gateup_swiglu.restype = c_i
        lib.asq_w8a8_rmsnorm_gateup_swiglu.argtypes = [c_vp, c_i, c_vp, c_f, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_f, c_f,
                                                       c_vp, c_f, c_i, c_vp, c_sz, c_vp]
        lib.asq_w8a8_linear_res.restype = c_i
        lib.asq_w8a8_linear_res.argtypes = [c_vp, c_i, c_vp, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_i, c_f, c_f,
                                            c_vp, c_vp, c_i, c_vp, c_sz, c_vp]
        lib.asq_w8a8_linear_q8_res.restype = c_i
        lib.asq_w8a8_linear_q8_res.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_f, c_vp, c_vp]
        lib.asq_w8a8_linear_q8_rope.restype = c_i
        lib.asq_w8a8_linear_q8_rope.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_f, c_vp,
                                                c_vp, c_vp, c_i64, c_i64, c_i64, c_i, c_vp]
        lib.asq_i8bmm.restype = c_i
        lib.asq_i8bmm.argtypes = [c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_i64, c_f, c_vp]
        lib.asq_w8a8_grouped_linear.restype = c_i
        lib.asq_w8a8_grouped_linear.argtypes = [c_vp, c_i, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_i, c_vp, c_vp, c_vp, c_vp,
                                                c_i, c_vp, c_i, c_i, c_vp, c_sz, c_vp]
        c_pp = ctypes.POINTER(ctypes.c_void_p)
        lib.asq_ar_buffer_bytes.restype = c_i
        lib.asq_ar_buffer_bytes.argtypes = [c_i64, c_i64, c_i, ctypes.POINTER(c_sz), ctypes.POINTER(c_sz)]
        lib.asq_w8a8_linear_q8_allreduce.restype = c_i
        lib.asq_w8a8_linear_q8_allreduce.argtypes = [c_vp, c_vp, c_vp, c_vp, c_pp, c_i, c_i64, c_i64, c_i64, c_f, c_vp,
                                                     c_pp, c_pp, c_i, c_i, c_i, c_vp, c_vp]
        lib.asq_q8_linear_allreduce_nvls.restype = c_i
        lib.asq_q8_linear_allreduce_nvls.argtypes = [c_vp, c_i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_f,
                                                     c_vp, c_pp, c_vp, c_vp, c_sz, c_i, c_i, c_i, c_vp]
        lib.asq_nvls_probe.restype = c_i
        lib.asq_nvls_probe.argtypes = [c_vp, c_vp, c_sz, c_i, c_i, c_i, c_i, c_vp, c_vp]
        lib.asq_dev_alloc.restype = c_i
        lib.asq_dev_alloc.argtypes = [c_sz, c_pp]
        lib.asq_dev_free.restype = c_i
        lib.asq_dev_free.argtypes = [c_vp]
        lib.asq_ipc_export.restype = c_i
        lib.asq_ipc_export.argtypes = [c_vp, c_vp]
        lib.asq_ipc_open.restype = c_i
        lib.asq_ipc_open.argtypes = [c_vp, c_pp]
        lib.asq_ipc_close.restype = c_i
        lib.asq_ipc_close.argtypes = [c_vp]
        lib.asq_add_rmsnorm_quant.restype = c_i
        lib.asq_add_rmsnorm_quant.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_f, c_vp]
        lib.asq_silu_mul_quant.restype = c_i
        lib.asq_silu_mul_quant.argtypes = [c_vp, c_i, c_i64, c_i64, c_i64, c_f, c_i, c_vp, c_vp, c_vp]
        lib.asq_rope_inplace.restype = c_i
        lib.asq_rope_inplace.argtypes = [c_vp, c_i, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp]
        _lib = lib
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        msg = load().asq_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libasq_b200 error {rc}: {msg}")


def _require_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "autosmoothquant_b200 kernels need CUDA tensors (sm_100a); there is no CPU fallback "
                f"(got a tensor on {t.device})"
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is None:
        raise RuntimeError("no tensor given")
    return dev


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _code(dtype: torch.dtype) -> int:
    try:
        return _DTYPE_CODE[dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {dtype}") from None


# ---------------------------------------------------------------- launch ordering across streams
# asq_linear_kernel is a persistent grid of one CTA (pair) per SM whose CTAs wait on each other inside the launch
# (phase-1 panel counters, the grid-wide absmax, stream-K partials, all-reduce flags).  That is live as long as every
# CTA of the grid eventually becomes resident: true next to any FINITE foreign kernel (NCCL, attention, copies — they
# drain and free their SMs), but NOT next to a second grid of the same kind on another stream: each could hold part of
# the SMs and spin on CTAs of its own that can never be scheduled.  So launches of one device are kept stream-ordered:
# when a launch arrives on a different stream than the previous one, the new stream first waits (event) for everything
# queued on the old one.  Launches on one stream — the normal case, and everything replayed from one CUDA graph — pay
# a dictionary look-up.  ASQ_STREAM_GUARD=0 removes the check (callers that order their streams themselves).  Users of
# the C ABI own this rule: see INTEGRATION.md, "Streams".
_STREAM_GUARD = os.environ.get("ASQ_STREAM_GUARD", "1") != "0"
_last_launch_stream: dict = {}  # device index -> torch.cuda.Stream that carried the most recent launch


def _order_after_previous_launch(dev: torch.device, cur) -> None:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    prev = _last_launch_stream.get(idx)
    if prev is not None and prev.cuda_stream == cur.cuda_stream:
        return
    if torch.cuda.is_current_stream_capturing():
        # inside a CUDA-graph capture: ordering against streams outside the capture is the capturing code's business
        # (torch.cuda.graph already makes the capture stream wait for the caller's stream), and an event from an
        # uncaptured stream would invalidate the capture.  The captured launches are ordered among themselves.
        return
    if prev is not None:
        try:
            cur.wait_stream(prev)
        except RuntimeError:
            pass  # `prev` is being captured by another thread: nothing of it runs now, nothing to wait for
    _last_launch_stream[idx] = cur


def _stream(dev: torch.device) -> int:
    cur = torch.cuda.current_stream(dev)
    if _STREAM_GUARD:
        try:
            _order_after_previous_launch(dev, cur)
        except Exception:  # noqa: BLE001 - bookkeeping must never be the reason a launch fails
            pass
    return cur.cuda_stream


# ---------------------------------------------------------------- workspace
# One scratch buffer per (device, stream): launches on a stream are ordered, so they can share it.
# It is zero-filled once (the kernels restore their phase counters before exiting) and only grows.
_ws_cache: dict = {}
_ws_lock = threading.Lock()


def _workspace(dev: torch.device, stream: int, nbytes: int) -> Tuple[int, int]:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), stream)
    with _ws_lock:
        buf = _ws_cache.get(key)
        if buf is None or buf.numel() < nbytes + 1024:
            size = max(nbytes + 1024, 1 << 25)
            size = 1 << (size - 1).bit_length()  # grow geometrically
            buf = torch.zeros(size, dtype=torch.uint8, device=dev)
            _ws_cache[key] = buf
    base = (buf.data_ptr() + 1023) & ~1023
    return base, buf.numel() - (base - buf.data_ptr())


def release_workspaces() -> None:
    with _ws_lock:
        _ws_cache.clear()


# ---------------------------------------------------------------- entry points
def w8a8_linear(
    x: torch.Tensor,
    weight: torch.Tensor,
    bias: Optional[torch.Tensor],
    act_mode: int,
    quant_scale: float = 1.0,
    dequant_scale: float = 1.0,
    col_scale: Optional[torch.Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    row_scale_out: Optional[torch.Tensor] = None,
    div_mode: Optional[int] = None,
    residual: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Fused quantise -> INT8 GEMM -> dequant (+bias) for a 2-D ``x`` [M,K]; returns [M,N].
    residual [M,N] (16-bit, the output dtype): returns T(residual + T(linear output)) from the same launch."""
    global _launches
    dev = _require_cuda(x, weight, bias, col_scale, row_scale_out, residual)
    if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1]:
        raise ValueError(f"shape mismatch: x {tuple(x.shape)} vs weight {tuple(weight.shape)}")
    if weight.dtype != torch.int8:
        raise TypeError(f"weight must be int8, got {weight.dtype}")
    if not x.is_contiguous():
        x = x.contiguous()
    if not weight.is_contiguous():
        raise ValueError("weight must be contiguous [N,K]")
    for name, t in (("bias", bias), ("col_scale", col_scale), ("row_scale_out", row_scale_out)):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise TypeError(f"{name} must be a contiguous float32 tensor")
    M, K = x.shape
    N = weight.shape[0]
    out_dtype = out_dtype or x.dtype
    y = torch.empty((M, N), dtype=out_dtype, device=dev)
    if M == 0:
        return y
    lib = load()
    with torch.cuda.device(dev):
        stream = _stream(dev)
        need = lib.asq_workspace_bytes(M, K)
        ws, ws_bytes = _workspace(dev, stream, need)
        if residual is not None:
            if residual.shape != (M, N) or residual.dtype != out_dtype or not residual.is_contiguous():
                raise ValueError("residual must be a contiguous [M,N] tensor of the output dtype")
            rc = lib.asq_w8a8_linear_res(
                x.data_ptr(), _code(x.dtype), weight.data_ptr(), _ptr(bias), residual.data_ptr(), y.data_ptr(), _code(out_dtype),
                M, N, K, act_mode, float(quant_scale), float(dequant_scale), _ptr(col_scale), _ptr(row_scale_out),
                _default_div_mode if div_mode is None else div_mode, ws, ws_bytes, stream,
            )
        else:
            rc = lib.asq_w8a8_linear(
                x.data_ptr(), _code(x.dtype), weight.data_ptr(), _ptr(bias), y.data_ptr(), _code(out_dtype),
                M, N, K, act_mode, float(quant_scale), float(dequant_scale), _ptr(col_scale), _ptr(row_scale_out),
                _default_div_mode if div_mode is None else div_mode, ws, ws_bytes, stream,
            )
    _check(rc)
    _launches += 1
    return y


def fp8_linear(
    x: torch.Tensor,
    weight: torch.Tensor,
    bias: Optional[torch.Tensor],
    act_mode: int,
    in_scale: float = 1.0,
    w_scale: float = 1.0,
    out_dtype: Optional[torch.dtype] = None,
    row_scale_out: Optional[torch.Tensor] = None,
    div_mode: Optional[int] = None,
    out_scale: float = 0.0,
    col_scale: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Fused quantise -> e4m3 GEMM (fp32 accumulate) -> scale (+bias); ``weight`` is float8_e4m3fn [N,K].
    col_scale [N] fp32 (device) replaces the scalar ``w_scale`` by one weight scale per output column (horizontally
    fused projections: asq_fp8_linear_cs); dynamic activation modes only."""
    global _launches
    dev = _require_cuda(x, weight, bias, row_scale_out, col_scale)
    if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1]:
        raise ValueError(f"shape mismatch: x {tuple(x.shape)} vs weight {tuple(weight.shape)}")
    if weight.dtype not in (torch.float8_e4m3fn, torch.uint8):
        raise TypeError(f"weight must be float8_e4m3fn, got {weight.dtype}")
    if not x.is_contiguous():
        x = x.contiguous()
    if not weight.is_contiguous():
        raise ValueError("weight must be contiguous [N,K]")
    if bias is not None and (bias.dtype != torch.float32 or not bias.is_contiguous()):
        raise TypeError("bias must be a contiguous float32 tensor")
    M, K = x.shape
    N = weight.shape[0]
    out_dtype = out_dtype or x.dtype
    y = torch.empty((M, N), dtype=out_dtype, device=dev)
    if M == 0:
        return y
    lib = load()
    with torch.cuda.device(dev):
        stream = _stream(dev)
        need = lib.asq_workspace_bytes(M, K)
        ws, ws_bytes = _workspace(dev, stream, need)
        if col_scale is not None:
            if col_scale.dtype != torch.float32 or col_scale.numel() != N or not col_scale.is_contiguous():
                raise ValueError("col_scale must be a contiguous float32 [N] tensor")
            if out_scale != 0.0 or act_mode == ACT_SCALE:
                raise ValueError("col_scale: dynamic activation scales only, no output fake-quantisation")
            rc = lib.asq_fp8_linear_cs(
                x.data_ptr(), _code(x.dtype), weight.data_ptr(), _ptr(bias), y.data_ptr(), _code(out_dtype),
                M, N, K, act_mode, col_scale.data_ptr(), _ptr(row_scale_out),
                _default_div_mode if div_mode is None else div_mode, ws, ws_bytes, stream,
            )
        else:
            rc = lib.asq_fp8_linear(
                x.data_ptr(), _code(x.dtype), weight.data_ptr(), _ptr(bias), y.data_ptr(), _code(out_dtype),
                M, N, K, act_mode, float(in_scale), float(w_scale), float(out_scale), _ptr(row_scale_out),
                _default_div_mode if div_mode is None else div_mode, ws, ws_bytes, stream,
            )
    _check(rc)
    _launches += 1
    return y


def i8gemm_o32(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor) -> None:
    """out[M,N] (int32) = a[M,K] (int8) @ w[N,K]^T (int8), exact, in place on ``out``."""
    global _launches
    dev = _require_cuda(a, w, out)
    if a.dtype != torch.int8 or w.dtype != torch.int8 or out.dtype != torch.int32:
        raise TypeError("i8gemm_o32 expects int8, int8, int32 tensors")
    if a.dim() != 2 or w.dim() != 2 or a.shape[1] != w.shape[1] or tuple(out.shape) != (a.shape[0], w.shape[0]):
        raise ValueError(f"shape mismatch: a {tuple(a.shape)} w {tuple(w.shape)} out {tuple(out.shape)}")
    if not (a.is_contiguous() and w.is_contiguous() and out.is_contiguous()):
        raise ValueError("i8gemm_o32 expects contiguous tensors")
    if a.shape[0] == 0:
        return
    lib = load()
    with torch.cuda.device(dev):
        stream = _stream(dev)
        ws, ws_bytes = _workspace(dev, stream, lib.asq_workspace_bytes(0, 0))
        rc = lib.asq_i8gemm_o32(a.data_ptr(), w.data_ptr(), out.data_ptr(), a.shape[0], w.shape[0], a.shape[1],
                                ws, ws_bytes, stream)
    _check(rc)
    _launches += 1


def i8gemm_epi(
    a: torch.Tensor,
    w: torch.Tensor,
    out: torch.Tensor,
    alpha: float,
    beta: float = 0.0,
    bias: Optional[torch.Tensor] = None,
    relu: bool = False,
) -> None:
    """out = convert(alpha * (a @ w^T) + beta * bias) with optional ReLU; ``out`` int8/int32/float."""
    global _launches
    dev = _require_cuda(a, w, out, bias)
    if a.dtype != torch.int8 or w.dtype != torch.int8:
        raise TypeError("i8gemm_epi expects int8 operands")
    if a.dim() != 2 or w.dim() != 2 or a.shape[1] != w.shape[1] or tuple(out.shape) != (a.shape[0], w.shape[0]):
        raise ValueError(f"shape mismatch: a {tuple(a.shape)} w {tuple(w.shape)} out {tuple(out.shape)}")
    if not (a.is_contiguous() and w.is_contiguous() and out.is_contiguous()):
        raise ValueError("i8gemm_epi expects contiguous tensors")
    if bias is not None and (not bias.is_contiguous() or bias.numel() != w.shape[0]):
        raise ValueError("bias must be a contiguous [N] tensor")
    if a.shape[0] == 0:
        return
    lib = load()
    with torch.cuda.device(dev):
        stream = _stream(dev)
        ws, ws_bytes = _workspace(dev, stream, lib.asq_workspace_bytes(0, 0))
        rc = lib.asq_i8gemm_epi(
            a.data_ptr(), w.data_ptr(), _ptr(bias), _code(bias.dtype) if bias is not None else ASQ_F32,
            out.data_ptr(), _code(out.dtype), a.shape[0], w.shape[0], a.shape[1], float(alpha), float(beta),
            EPI_RELU if relu else 0, ws, ws_bytes, stream,
        )
    _check(rc)
    _launches += 1


def quantize_act(
    x: torch.Tensor,
    act_mode: int,
    quant_scale: float = 1.0,
    fp8: bool = False,
    div_mode: Optional[int] = None,
    row_scale: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Parity tap of the fused prologue: returns (q [M,K] int8 | float8_e4m3fn, row_scale [M] or None)."""
    global _launches
    dev = _require_cuda(x)
    if x.dim() != 2:
        raise ValueError("quantize_act expects a 2-D tensor")
    if not x.is_contiguous():
        x = x.contiguous()
    M, K = x.shape
    q = torch.empty((M, K), dtype=torch.uint8 if fp8 else torch.int8, device=dev)
    if act_mode == ACT_ROW_SCALE_GIVEN:
        if row_scale is None or row_scale.dtype != torch.float32 or row_scale.numel() != M:
            raise ValueError("ACT_ROW_SCALE_GIVEN needs row_scale: float32 [M]")
        rs = row_scale.contiguous()
    else:
        rs = torch.empty((M,), dtype=torch.float32, device=dev) if act_mode == ACT_PER_TOKEN else None
    if M > 0:
        with torch.cuda.device(dev):
            rc = load().asq_quantize_act(
                x.data_ptr(), _code(x.dtype), q.data_ptr(), _ptr(rs), M, K, act_mode, float(quant_scale),
                _default_div_mode if div_mode is None else div_mode, 1 if fp8 else 0, _stream(dev),
            )
        _check(rc)
        _launches += 1
    if fp8:
        q = q.view(torch.float8_e4m3fn)
    return q, rs


# ---------------------------------------------------------------- producer-side fusions (asq_glue.cu)
def w8a8_linear_q8(
    xq: torch.Tensor,
    weight: torch.Tensor,
    bias: Optional[torch.Tensor],
    dequant_scale: float = 1.0,
    col_scale: Optional[torch.Tensor] = None,
    row_scale: Optional[torch.Tensor] = None,
    out_dtype: torch.dtype = torch.bfloat16,
    rope: Optional[tuple] = None,
    residual: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """INT8 GEMM + dequant epilogue for activations a fused producer already quantised (no prologue).
    rope = (cos, sin, S, rope_cols): rotate-half RoPE on columns < rope_cols in the epilogue; cos / sin are the
    [S, 128] tables in the blocked layout of `rope_tables_blocked`; an optional fifth element True vouches that the
    tables repeat their first half (HF's cat(freqs, freqs)), which halves the table reads."""
    global _launches
    dev = _require_cuda(xq, weight, bias, col_scale, row_scale)
    if xq.dtype != torch.int8 or weight.dtype != torch.int8 or xq.dim() != 2 or xq.shape[1] != weight.shape[1]:
        raise ValueError("w8a8_linear_q8 expects int8 [M,K] activations and int8 [N,K] weights")
    if not (xq.is_contiguous() and weight.is_contiguous()):
        raise ValueError("w8a8_linear_q8 expects contiguous tensors")
    M, K = xq.shape
    N = weight.shape[0]
    y = torch.empty((M, N), dtype=out_dtype, device=dev)
    if M == 0:
        return y
    lib = load()
    if residual is not None:
        _require_cuda(residual)
        if rope is not None or residual.shape != (M, N) or residual.dtype != out_dtype or not residual.is_contiguous():
            raise ValueError("residual must be a contiguous [M,N] tensor of the output dtype (not combinable with rope)")
        with torch.cuda.device(dev):
            rc = lib.asq_w8a8_linear_q8_res(xq.data_ptr(), _ptr(row_scale), weight.data_ptr(), _ptr(bias), residual.data_ptr(),
                                            y.data_ptr(), _code(out_dtype), M, N, K, float(dequant_scale), _ptr(col_scale),
                                            _stream(dev))
        _check(rc)
        _launches += 1
        return y
    if rope is not None:
        cos, sin, S, rope_cols = rope[:4]
        halves_equal = bool(rope[4]) if len(rope) > 4 else False
        _require_cuda(cos, sin)
        if cos.dtype != out_dtype or sin.dtype != out_dtype or cos.shape != sin.shape or cos.dim() != 3 or not (
                cos.is_contiguous() and sin.is_contiguous()) or cos.shape[1] != S or cos.shape[2] != 8:
            raise ValueError("rope tables must be rope_tables_blocked() outputs [head_dim/8, S, 8] of the output dtype")
        with torch.cuda.device(dev):
            rc = lib.asq_w8a8_linear_q8_rope(xq.data_ptr(), _ptr(row_scale), weight.data_ptr(), _ptr(bias), y.data_ptr(),
                                             _code(out_dtype), M, N, K, float(dequant_scale), _ptr(col_scale),
                                             cos.data_ptr(), sin.data_ptr(), int(S), int(rope_cols), cos.shape[0] * 8,
                                             1 if halves_equal else 0, _stream(dev))
        _check(rc)
        _launches += 1
        return y
    with torch.cuda.device(dev):
        stream = _stream(dev)
        ws, ws_bytes = _workspace(dev, stream, lib.asq_workspace_bytes(0, 0))
        rc = lib.asq_w8a8_linear_q8(xq.data_ptr(), _ptr(row_scale), weight.data_ptr(), _ptr(bias), y.data_ptr(),
                                    _code(out_dtype), M, N, K, float(dequant_scale), _ptr(col_scale), ws, ws_bytes, stream)
    _check(rc)
    _launches += 1
    return y


def rope_tables_blocked(table: torch.Tensor) -> torch.Tensor:
    """[S, head_dim] cos or sin table -> the [head_dim/8, S, 8] layout the RoPE epilogue reads (one-time re-layout)."""
    S, hd = table.shape
    return table.reshape(S, hd // 8, 8).transpose(0, 1).contiguous()


def w8a8_rmsnorm_linear(
    x: torch.Tensor,
    norm_weight: torch.Tensor,
    eps: float,
    weight: torch.Tensor,
    bias: Optional[torch.Tensor],
    dequant_scale: float = 1.0,
    col_scale: Optional[torch.Tensor] = None,
    rope: Optional[tuple] = None,
) -> torch.Tensor:
    """RMSNorm (folded scale) -> round-only int8 -> GEMM -> dequant [-> RoPE] in ONE launch; x [M,K] f16 | bf16.
    rope as in w8a8_linear_q8.  Bit-identical to add_rmsnorm_quant(x, None, ...) followed by w8a8_linear_q8."""
    global _launches
    dev = _require_cuda(x, norm_weight, weight, bias, col_scale)
    if x.dim() != 2 or weight.dtype != torch.int8 or x.shape[1] != weight.shape[1] or norm_weight.dtype != x.dtype:
        raise ValueError("w8a8_rmsnorm_linear expects x [M,K] (16-bit), a [K] norm weight of the same dtype and int8 [N,K] weights")
    if not (x.is_contiguous() and weight.is_contiguous() and norm_weight.is_contiguous()):
        raise ValueError("w8a8_rmsnorm_linear expects contiguous tensors")
    M, K = x.shape
    N = weight.shape[0]
    y = torch.empty((M, N), dtype=x.dtype, device=dev)
    if M == 0:
        return y
    cos = sin = None
    S = rope_cols = hd = 0
    dup = False
    if rope is not None:
        cos, sin, S, rope_cols = rope[:4]
        dup = bool(rope[4]) if len(rope) > 4 else False
        _require_cuda(cos, sin)
        if cos.dtype != x.dtype or sin.dtype != x.dtype or cos.shape != sin.shape or cos.dim() != 3 or cos.shape[1] != S or cos.shape[2] != 8:
            raise ValueError("rope tables must be rope_tables_blocked() outputs [head_dim/8, S, 8] of the activation dtype")
        hd = cos.shape[0] * 8
    lib = load()
    with torch.cuda.device(dev):
        stream = _stream(dev)
        ws, ws_bytes = _workspace(dev, stream, lib.asq_workspace_bytes(M, K))
        rc = lib.asq_w8a8_rmsnorm_linear_rope(x.data_ptr(), _code(x.dtype), norm_weight.data_ptr(), float(eps), weight.data_ptr(),
                                              _ptr(bias), y.data_ptr(), M, N, K, float(dequant_scale), _ptr(col_scale),
                                              _ptr(cos), _ptr(sin), int(S), int(rope_cols), int(hd), 1 if dup else 0, ws, ws_bytes, stream)
    _check(rc)
    _launches += 1
    return y


def w8a8_rmsnorm_gateup_swiglu(
    x: torch.Tensor,
    norm_weight: torch.Tensor,
    eps: float,
    weight_il: torch.Tensor,
    bias_il: Optional[torch.Tensor],
    dequant_scale: float = 1.0,
    up_dequant_scale: Optional[float] = None,
    col_scale_il: Optional[torch.Tensor] = None,
    out_quant_scale: Optional[float] = None,
    div_mode: Optional[int] = None,
) -> torch.Tensor:
    """RMSNorm -> int8 -> gate|up GEMM -> SwiGLU (-> down_proj's int8) in ONE launch; see w8a8_gateup_swiglu."""
    global _launches
    dev = _require_cuda(x, norm_weight, weight_il, bias_il, col_scale_il)
    if x.dim() != 2 or weight_il.dtype != torch.int8 or x.shape[1] != weight_il.shape[1] or norm_weight.dtype != x.dtype:
        raise ValueError("w8a8_rmsnorm_gateup_swiglu expects x [M,K] (16-bit), a [K] norm weight and int8 [2I,K] weights")
    if not (x.is_contiguous() and weight_il.is_contiguous() and norm_weight.is_contiguous()):
        raise ValueError("w8a8_rmsnorm_gateup_swiglu expects contiguous tensors")
    M, K = x.shape
    N = weight_il.shape[0]
    out_dtype = torch.int8 if out_quant_scale is not None else x.dtype
    out = torch.empty((M, N // 2), dtype=out_dtype, device=dev)
    if M == 0:
        return out
    lib = load()
    with torch.cuda.device(dev):
        stream = _stream(dev)
        ws, ws_bytes = _workspace(dev, stream, lib.asq_workspace_bytes(M, K))
        rc = lib.asq_w8a8_rmsnorm_gateup_swiglu(
            x.data_ptr(), _code(x.dtype), norm_weight.data_ptr(), float(eps), weight_il.data_ptr(), _ptr(bias_il), out.data_ptr(),
            _code(out_dtype), M, N, K, float(dequant_scale),
            float(dequant_scale if up_dequant_scale is None else up_dequant_scale), _ptr(col_scale_il),
            float(out_quant_scale) if out_quant_scale is not None else 0.0,
            _default_div_mode if div_mode is None else div_mode, ws, ws_bytes, stream)
    _check(rc)
    _launches += 1
    return out


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """Load-time re-layout for w8a8_gateup_swiglu: blocks of 32 gate rows (or vector entries) alternate with
    the 32 matching up rows, so one 64-column accumulator group holds both operands of 32 SwiGLU outputs."""
    if gate.shape != up.shape or gate.shape[0] % 32:
        raise ValueError("gate / up must have the same shape with a leading dimension that is a multiple of 32")
    I = gate.shape[0]
    rest = gate.shape[1:]
    return torch.stack((gate.reshape(I // 32, 32, *rest), up.reshape(I // 32, 32, *rest)), dim=1).reshape(2 * I, *rest).contiguous()


def w8a8_gateup_swiglu(
    xq: torch.Tensor,
    weight_il: torch.Tensor,
    bias_il: Optional[torch.Tensor],
    dequant_scale: float = 1.0,
    col_scale_il: Optional[torch.Tensor] = None,
    row_scale: Optional[torch.Tensor] = None,
    out_quant_scale: Optional[float] = None,
    up_dequant_scale: Optional[float] = None,
    mid_dtype: torch.dtype = torch.bfloat16,
    div_mode: Optional[int] = None,
) -> torch.Tensor:
    """gate|up INT8 GEMM whose epilogue applies SiLU(gate) * up and (out_quant_scale given) the per-tensor
    quantisation of the next Linear: returns int8 [M, I], or the product in mid_dtype when out_quant_scale is None.
    weight_il / bias_il / col_scale_il are in interleave_gate_up order; dequant_scale applies to the gate columns,
    up_dequant_scale (default: the same) to the up columns, col_scale_il overrides both per column."""
    global _launches
    dev = _require_cuda(xq, weight_il, bias_il, col_scale_il, row_scale)
    if xq.dtype != torch.int8 or weight_il.dtype != torch.int8 or xq.dim() != 2 or xq.shape[1] != weight_il.shape[1]:
        raise ValueError("w8a8_gateup_swiglu expects int8 [M,K] activations and int8 [2I,K] weights")
    if not (xq.is_contiguous() and weight_il.is_contiguous()):
        raise ValueError("w8a8_gateup_swiglu expects contiguous tensors")
    M, K = xq.shape
    N = weight_il.shape[0]
    out_dtype = torch.int8 if out_quant_scale is not None else mid_dtype
    out = torch.empty((M, N // 2), dtype=out_dtype, device=dev)
    if M == 0:
        return out
    with torch.cuda.device(dev):
        rc = load().asq_w8a8_gateup_swiglu_q8(
            xq.data_ptr(), _ptr(row_scale), weight_il.data_ptr(), _ptr(bias_il), out.data_ptr(), _code(out_dtype),
            _code(mid_dtype), M, N, K, float(dequant_scale),
            float(dequant_scale if up_dequant_scale is None else up_dequant_scale), _ptr(col_scale_il),
            float(out_quant_scale) if out_quant_scale is not None else 0.0,
            _default_div_mode if div_mode is None else div_mode, _stream(dev))
    _check(rc)
    _launches += 1
    return out


def i8bmm(a: torch.Tensor, b: torch.Tensor, out_dtype: torch.dtype, alpha: float = 1.0) -> torch.Tensor:
    """c[B,M,N] = epilogue(alpha * a[B,M,K] @ b[B,N,K]^T): int32 raw, float32 scaled, or int8 rounded + saturated."""
    global _launches
    dev = _require_cuda(a, b)
    if a.dtype != torch.int8 or b.dtype != torch.int8 or a.dim() != 3 or b.dim() != 3 or a.shape[0] != b.shape[0] or a.shape[2] != b.shape[2]:
        raise ValueError("i8bmm expects int8 a [B,M,K] and b [B,N,K]")
    if out_dtype not in (torch.int8, torch.int32, torch.float32):
        raise ValueError("i8bmm output dtype must be int8, int32 or float32")
    a, b = a.contiguous(), b.contiguous()
    B, M, K = a.shape
    N = b.shape[1]
    c = torch.empty((B, M, N), dtype=out_dtype, device=dev)
    if B * M * N == 0:
        return c
    with torch.cuda.device(dev):
        rc = load().asq_i8bmm(a.data_ptr(), b.data_ptr(), c.data_ptr(), _code(out_dtype), B, M, N, K, float(alpha), _stream(dev))
    _check(rc)
    _launches += 1
    return c


def w8a8_grouped_linear(
    x: torch.Tensor,
    weight_stacked: torch.Tensor,
    group_of_blk: torch.Tensor,
    group_dequant_scale: torch.Tensor,
    act_mode: int,
    group_quant_scale: Optional[torch.Tensor] = None,
    group_dequant_scale_up: Optional[torch.Tensor] = None,
    swiglu: bool = False,
    div_mode: Optional[int] = None,
    row_scale: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """All experts of an MoE block in one launch (include/asq.h: asq_w8a8_grouped_linear).  x [M_pad, K] holds the
    routed rows sorted by expert, every expert's segment padded to a multiple of 256 rows; group_of_blk (int32
    [M_pad / 128]) names the expert of each 128-row block (-1: unused); weight_stacked [G * N, K].
    act_mode ACT_ROW_SCALE_GIVEN quantises with the caller's row_scale [M_pad] fp32 (tensor-parallel w2)."""
    global _launches
    dev = _require_cuda(x, weight_stacked, group_of_blk, group_dequant_scale, group_quant_scale, group_dequant_scale_up, row_scale)
    if act_mode == ACT_ROW_SCALE_GIVEN and (row_scale is None or row_scale.dtype != torch.float32 or row_scale.numel() != x.shape[0]
                                            or not row_scale.is_contiguous()):
        raise ValueError("ACT_ROW_SCALE_GIVEN needs row_scale: contiguous float32 [M_pad]")
    G = group_dequant_scale.numel()
    if x.dim() != 2 or weight_stacked.dtype != torch.int8 or weight_stacked.shape[1] != x.shape[1] or weight_stacked.shape[0] % G:
        raise ValueError("w8a8_grouped_linear expects x [M_pad,K] and stacked int8 weights [G*N,K]")
    if group_of_blk.dtype != torch.int32 or group_of_blk.numel() * 128 < x.shape[0]:
        raise ValueError("group_of_blk must be int32 with one entry per 128 rows of x")
    if not (x.is_contiguous() and weight_stacked.is_contiguous() and group_of_blk.is_contiguous()):
        raise ValueError("w8a8_grouped_linear expects contiguous tensors")
    M, K = x.shape
    N = weight_stacked.shape[0] // G
    out_dtype = x.dtype
    y = torch.empty((M, N // 2 if swiglu else N), dtype=out_dtype, device=dev)
    if M == 0:
        return y
    lib = load()
    with torch.cuda.device(dev):
        stream = _stream(dev)
        ws, ws_bytes = _workspace(dev, stream, lib.asq_workspace_bytes(M, K))
        rc = lib.asq_w8a8_grouped_linear(
            x.data_ptr(), _code(x.dtype), weight_stacked.data_ptr(), y.data_ptr(), _code(out_dtype), M, N, K, G,
            group_of_blk.data_ptr(), group_dequant_scale.data_ptr(), _ptr(group_dequant_scale_up), _ptr(group_quant_scale),
            act_mode, _ptr(row_scale) if act_mode == ACT_ROW_SCALE_GIVEN else None, 1 if swiglu else 0,
            _default_div_mode if div_mode is None else div_mode, ws, ws_bytes, stream)
    _check(rc)
    _launches += 1
    return y


def add_rmsnorm_quant(
    x: torch.Tensor,
    delta: Optional[torch.Tensor],
    weight: torch.Tensor,
    eps: float,
    want_h: bool = False,
    want_q: bool = True,
) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """(x + delta, rmsnorm(x + delta) * weight in x's dtype, its int8 rounding); x is [M,H], updated out of place."""
    global _launches
    dev = _require_cuda(x, delta, weight)
    if x.dim() != 2 or not x.is_contiguous() or (delta is not None and (delta.shape != x.shape or not delta.is_contiguous())):
        raise ValueError("add_rmsnorm_quant expects contiguous [M,H] tensors")
    M, H = x.shape
    x_out = torch.empty_like(x) if delta is not None else x
    h = torch.empty_like(x) if want_h else None
    q = torch.empty((M, H), dtype=torch.int8, device=dev) if want_q else None
    if M > 0:
        with torch.cuda.device(dev):
            rc = load().asq_add_rmsnorm_quant(x.data_ptr(), _ptr(delta), weight.data_ptr(),
                                              x_out.data_ptr() if delta is not None else None, _ptr(h), _ptr(q),
                                              _code(x.dtype), M, H, float(eps), _stream(dev))
        _check(rc)
        _launches += 1
    return x_out, h, q


def silu_mul_quant(
    gate_up: torch.Tensor,
    quant_scale: float = 1.0,
    want_q: bool = True,
    want_a: bool = False,
    div_mode: Optional[int] = None,
) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """gate_up [M, 2I] -> (int8 sat(rint(T(a / quant_scale))), a = T(T(silu(gate)) * up)), each optional."""
    global _launches
    dev = _require_cuda(gate_up)
    if gate_up.dim() != 2 or gate_up.shape[1] % 2 or gate_up.stride(1) != 1:
        raise ValueError("silu_mul_quant expects a [M, 2I] tensor with unit inner stride")
    M, I = gate_up.shape[0], gate_up.shape[1] // 2
    q = torch.empty((M, I), dtype=torch.int8, device=dev) if want_q else None
    a = torch.empty((M, I), dtype=gate_up.dtype, device=dev) if want_a else None
    if M > 0:
        with torch.cuda.device(dev):
            rc = load().asq_silu_mul_quant(gate_up.data_ptr(), _code(gate_up.dtype), M, I, gate_up.stride(0),
                                           float(quant_scale), _default_div_mode if div_mode is None else div_mode,
                                           _ptr(q), _ptr(a), _stream(dev))
        _check(rc)
        _launches += 1
    return q, a


def rope_inplace(qk: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, seq_len: int, n_heads: int, head_dim: int) -> None:
    """Rotate-half RoPE in place on the first n_heads*head_dim columns of qk [M, row]; position = row % seq_len."""
    global _launches
    dev = _require_cuda(qk, cos, sin)
    if qk.dim() != 2 or qk.stride(1) != 1 or cos.dtype != qk.dtype or sin.dtype != qk.dtype:
        raise ValueError("rope_inplace expects [M, row] activations and tables of the same dtype")
    if tuple(cos.shape) != (seq_len, head_dim) or tuple(sin.shape) != (seq_len, head_dim) or not (cos.is_contiguous() and sin.is_contiguous()):
        raise ValueError("cos/sin tables must be contiguous [seq_len, head_dim]")
    if qk.shape[0] == 0:
        return
    with torch.cuda.device(dev):
        rc = load().asq_rope_inplace(qk.data_ptr(), _code(qk.dtype), cos.data_ptr(), sin.data_ptr(), qk.shape[0], seq_len,
                                     qk.stride(0), n_heads, head_dim, _stream(dev))
    _check(rc)
    _launches += 1
