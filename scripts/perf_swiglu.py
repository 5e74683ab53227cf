"""Epilogue cost sweep on the gate|up shape and o_proj prologue alternatives: python scripts/perf_swiglu.py"""
import sys
import torch
sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L
dev = torch.device("cuda:0")

def timeit(fn, iters=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e6

M, I, K = 2048, 11008, 4096
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
w = torch.randint(-128, 128, (2 * I, K), dtype=torch.int8, device=dev)
cs = torch.full((2 * I,), 3e-6, device=dev)
o32 = torch.empty((M, 2 * I), dtype=torch.int32, device=dev)
print("gate|up 2048x22016x4096")
print(" i8gemm_o32          %7.1f us" % timeit(lambda: L.i8gemm_o32(a, w, o32)))
print(" q8 -> bf16 (cs)     %7.1f us" % timeit(lambda: L.w8a8_linear_q8(a, w, None, 1.0, col_scale=cs)))
print(" q8 -> bf16 (scalar) %7.1f us" % timeit(lambda: L.w8a8_linear_q8(a, w, None, 3e-6)))
print(" swiglu -> int8 (cs) %7.1f us" % timeit(lambda: L.w8a8_gateup_swiglu(a, w, None, 1.0, col_scale_il=cs, out_quant_scale=0.05)))
print(" swiglu -> int8 (sc) %7.1f us" % timeit(lambda: L.w8a8_gateup_swiglu(a, w, None, 3e-6, out_quant_scale=0.05)))
print(" swiglu -> bf16 (cs) %7.1f us" % timeit(lambda: L.w8a8_gateup_swiglu(a, w, None, 1.0, col_scale_il=cs)))
gu = L.w8a8_linear_q8(a, w, None, 1.0, col_scale=cs)
print(" silu_mul_quant      %7.1f us" % timeit(lambda: L.silu_mul_quant(gu, 0.05)))

M, N, K = 2048, 4096, 4096
x = (torch.randn(M, K, device=dev)).to(torch.bfloat16)
w = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
print("o_proj 2048x4096x4096")
print(" fused ACT_SCALE     %7.1f us" % timeit(lambda: L.w8a8_linear(x, w, None, L.ACT_SCALE, 0.05, 0.003)))
print(" fused ACT_ROUND     %7.1f us" % timeit(lambda: L.w8a8_linear(x, w, None, L.ACT_ROUND, 0.05, 0.003)))
print(" fused PER_TOKEN     %7.1f us" % timeit(lambda: L.w8a8_linear(x, w, None, L.ACT_PER_TOKEN, 0.05, 0.003)))
print(" q8 only             %7.1f us" % timeit(lambda: L.w8a8_linear_q8(a, w, None, 0.003)))
print(" quantize_act        %7.1f us" % timeit(lambda: L.quantize_act(x, L.ACT_SCALE, 0.05)))
def two():
    q, _ = L.quantize_act(x, L.ACT_SCALE, 0.05)
    return L.w8a8_linear_q8(q, w, None, 0.003)
print(" quantize + q8       %7.1f us" % timeit(two))
