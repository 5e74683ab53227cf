"""stream-K tail on/off for the plain-epilogue launches (run with ASQ_STREAMK=1|2): python scripts/perf_sk.py"""
import sys
import torch
sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L
from scripts.perf_epi_k import timeit
dev = torch.device("cuda:0")
for (M, N, K) in [(2048, 4096, 4096), (2048, 4096, 11008), (2048, 12288, 4096), (2048, 11008, 4096)]:
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(3)]
    i = [0]
    def q8():
        i[0] += 1
        return L.w8a8_linear_q8(a, ws[i[0] % 3], None, 3e-6)
    def fused():
        i[0] += 1
        return L.w8a8_linear(x, ws[i[0] % 3], None, L.ACT_SCALE, 0.05, 3e-6)
    print(f"{M}x{N}x{K}: q8 {timeit(q8, iters=30):7.1f} us | fused {timeit(fused, iters=30):7.1f} us", flush=True)
