"""GEMM-only perf sweep: python scripts/perf_gemm.py [fused]"""
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
    return e0.elapsed_time(e1) / iters * 1e-3

shapes = [(2048, 4096, 4096), (2048, 12288, 4096), (2048, 11008, 4096), (2048, 22016, 4096), (2048, 4096, 11008), (8192, 8192, 8192)]
fused = len(sys.argv) > 1 and sys.argv[1] == "fused"
line = []
for (M, N, K) in shapes:
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    w = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    out = torch.empty((M, N), dtype=torch.int32, device=dev)
    x = (torch.randn(M, K, device=dev) * 40).to(torch.bfloat16)
    t = timeit(lambda: L.i8gemm_o32(a, w, out))
    s = f"{M}x{N}x{K}: o32 {t*1e6:7.1f}us {2.0*M*N*K/t/1e12:6.0f}T"
    if fused:
        t2 = timeit(lambda: L.w8a8_linear(x, w, None, L.ACT_ROUND, 0.05, 0.003))
        t3 = timeit(lambda: L.w8a8_linear(x, w, None, L.ACT_PER_TOKEN, 0.05, 0.003))
        s += f" | round {t2*1e6:7.1f}us | token {t3*1e6:7.1f}us"
    line.append(s)
print("\n".join(line), flush=True)
