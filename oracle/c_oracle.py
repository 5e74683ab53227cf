"""ctypes loader for the plain-C oracle (oracle/asq_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "asq_oracle.c"
LIB = HERE / "_build" / "libasq_oracle.so"


def build(force: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(LIB), str(SRC), "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed: {res.stderr}")
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def i8gemm_o32(a: np.ndarray, w: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, np.int8)
    w = np.ascontiguousarray(w, np.int8)
    c = np.empty((a.shape[0], w.shape[0]), np.int32)
    load().oracle_i8gemm_o32(_p(a), _p(w), _p(c), ctypes.c_int64(a.shape[0]), ctypes.c_int64(w.shape[0]),
                             ctypes.c_int64(a.shape[1]))
    return c


def quant_per_token_f32(x: np.ndarray):
    x = np.ascontiguousarray(x, np.float32)
    q = np.empty(x.shape, np.int8)
    s = np.empty(x.shape[0], np.float32)
    load().oracle_quant_per_token_f32(_p(x), _p(q), _p(s), ctypes.c_int64(x.shape[0]), ctypes.c_int64(x.shape[1]))
    return q, s


def quant_round_f32(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    q = np.empty(x.shape, np.int8)
    load().oracle_quant_round_f32(_p(x), _p(q), ctypes.c_int64(x.size))
    return q


def dequant_f32(acc: np.ndarray, row_scale, ds: float, bias) -> np.ndarray:
    acc = np.ascontiguousarray(acc, np.int32)
    y = np.empty(acc.shape, np.float32)
    rs = None if row_scale is None else np.ascontiguousarray(row_scale, np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    load().oracle_dequant_f32(_p(acc), None if rs is None else _p(rs), ctypes.c_float(ds), None if b is None else _p(b),
                              _p(y), ctypes.c_int64(acc.shape[0]), ctypes.c_int64(acc.shape[1]))
    return y


def e4m3_encode(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    q = np.empty(x.shape, np.uint8)
    load().oracle_e4m3_encode(_p(x), _p(q), ctypes.c_int64(x.size))
    return q


def e4m3_decode(q: np.ndarray) -> np.ndarray:
    q = np.ascontiguousarray(q, np.uint8)
    x = np.empty(q.shape, np.float32)
    load().oracle_e4m3_decode(_p(q), _p(x), ctypes.c_int64(q.size))
    return x


def fp8_quant_per_token_f32(x: np.ndarray):
    x = np.ascontiguousarray(x, np.float32)
    q = np.empty(x.shape, np.uint8)
    s = np.empty(x.shape[0], np.float32)
    load().oracle_fp8_quant_per_token_f32(_p(x), _p(q), _p(s), ctypes.c_int64(x.shape[0]), ctypes.c_int64(x.shape[1]))
    return q, s


def fp8_linear_f64(a: np.ndarray, w: np.ndarray, a_scale: np.ndarray, w_scale: float, bias=None) -> np.ndarray:
    a = np.ascontiguousarray(a, np.uint8)
    w = np.ascontiguousarray(w, np.uint8)
    s = np.ascontiguousarray(a_scale, np.float32).reshape(-1)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    y = np.empty((a.shape[0], w.shape[0]), np.float64)
    load().oracle_fp8_linear_f64(_p(a), _p(w), _p(s), ctypes.c_float(w_scale), None if b is None else _p(b), _p(y),
                                 ctypes.c_int64(a.shape[0]), ctypes.c_int64(w.shape[0]), ctypes.c_int64(a.shape[1]))
    return y
