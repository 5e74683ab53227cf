"""Decode-sized M: GPU time per launch from CUDA-graph replays (host launch overhead excluded)."""
import sys, torch
sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L
dev = torch.device("cuda:0")
REPS = 24
def graph_time(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * REPS) * 1e3
for M in (1, 16, 64, 128, 256, 512):
    line = f"M={M:4d}:"
    for (N, K) in ((4096, 4096), (12288, 4096), (22016, 4096), (4096, 11008)):
        ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(8)]  # > L2 in total
        x = (torch.randn(M, K, device=dev) * 40).to(torch.bfloat16)
        i = [0]
        def f():
            i[0] += 1
            return L.w8a8_linear(x, ws[i[0] % 8], None, L.ACT_ROUND, 1.0, 0.003)
        t = graph_time(f)
        line += f"  {N}x{K}: {t:6.1f}us {N*K/t/1e3:5.0f}GB/s"
    print(line, flush=True)
