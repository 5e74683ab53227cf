"""Epilogue-vs-main-loop fit: time per launch as K varies (N = 22016, M = 2048): python scripts/perf_epi_k.py"""
import sys
import torch
sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L
dev = torch.device("cuda:0")

def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def main():
    M, I = 2048, 11008
    cs = torch.full((2 * I,), 3e-6, device=dev)
    bias = torch.zeros((2 * I,), device=dev)
    o32 = torch.empty((M, 2 * I), dtype=torch.int32, device=dev)
    print("K     | o32   | bf16 sc | bf16 cs | bf16 cs+b | swi8 sc | swi8 cs | swi16 cs  (us per launch)")
    for K in (256, 1024, 2048, 4096, 8192):
        a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
        w = torch.randint(-128, 128, (2 * I, K), dtype=torch.int8, device=dev)
        r = [timeit(lambda: L.i8gemm_o32(a, w, o32)),
             timeit(lambda: L.w8a8_linear_q8(a, w, None, 3e-6)),
             timeit(lambda: L.w8a8_linear_q8(a, w, None, 1.0, col_scale=cs)),
             timeit(lambda: L.w8a8_linear_q8(a, w, bias, 1.0, col_scale=cs)),
             timeit(lambda: L.w8a8_gateup_swiglu(a, w, None, 3e-6, out_quant_scale=0.05)),
             timeit(lambda: L.w8a8_gateup_swiglu(a, w, None, 1.0, col_scale_il=cs, out_quant_scale=0.05)),
             timeit(lambda: L.w8a8_gateup_swiglu(a, w, None, 1.0, col_scale_il=cs))]
        print(f"{K:5d} | " + " | ".join(f"{v:7.1f}" for v in r), flush=True)


if __name__ == "__main__":
    main()
