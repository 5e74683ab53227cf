"""Tiny launch sequence for ncu captures: usage python scripts/prof_target.py M N K [reps]."""
import sys
import torch
sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (2048, 4096, 4096)
dev = torch.device("cuda:0")
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
w = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
out = torch.empty((M, N), dtype=torch.int32, device=dev)
x = (torch.randn(M, K, device=dev) * 40).to(torch.bfloat16)
for _ in range(3):
    L.i8gemm_o32(a, w, out)
L.w8a8_linear(x, w, None, L.ACT_ROUND, 0.05, 0.003)
L.w8a8_linear(x, w, None, L.ACT_PER_TOKEN, 0.05, 0.003)
torch.cuda.synchronize()
