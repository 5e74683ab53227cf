"""ncu target: module forwards at M = 16 for the four Llama-2-7B projection shapes (weight-streaming kernel)."""
import sys
import torch
sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L
dev = torch.device("cuda:0")
for (N, K) in ((4096, 4096), (12288, 4096), (22016, 4096), (4096, 11008)):
    w = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    x = (torch.randn(16, K, device=dev) * 40).to(torch.bfloat16)
    for _ in range(2):
        L.w8a8_linear(x, w, None, L.ACT_ROUND, 1.0, 0.003)
torch.cuda.synchronize()
