"""The four GEMM launches of a Llama-2-7B layer (int8 in, bf16 out), graph-replayed: python scripts/perf_layer_shapes.py"""
import sys
import torch
sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L
from scripts.perf_epi_k import timeit
dev = torch.device("cuda:0")
for (M, N, K) in [(2048, 12288, 4096), (2048, 4096, 4096), (2048, 22016, 4096), (2048, 4096, 11008)]:
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(3)]
    i = [0]
    def run():
        i[0] += 1
        return L.w8a8_linear_q8(a, ws[i[0] % 3], None, 3e-6)
    t = timeit(run, iters=30)
    print(f"{M}x{N}x{K}: {t:7.1f} us  {2.0 * M * N * K / t / 1e6:6.0f} TOPS", flush=True)
