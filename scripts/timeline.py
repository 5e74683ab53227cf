"""Per-CTA phase timeline of one fused launch: python scripts/timeline.py M N K [mode]"""
import os, sys
import torch
sys.path.insert(0, ".")
dev = torch.device("cuda:0")
dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
os.environ["ASQ_DEBUG_TIMELINE"] = hex(dbg.data_ptr())
from autosmoothquant_b200 import _lib as L
M, N, K = (int(v) for v in sys.argv[1:4])
mode = {"round": L.ACT_ROUND, "scale": L.ACT_SCALE, "token": L.ACT_PER_TOKEN, "o32": -1}[sys.argv[4] if len(sys.argv) > 4 else "round"]
w = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
x = (torch.randn(M, K, device=dev) * 40).to(torch.bfloat16)
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
out = torch.empty((M, N), dtype=torch.int32, device=dev)
def run():
    if mode < 0:
        L.i8gemm_o32(a, w, out)
    else:
        L.w8a8_linear(x, w, None, mode, 0.05, 0.003)
for _ in range(5):
    run()
torch.cuda.synchronize()
dbg.zero_()
torch.cuda.synchronize()
run()
torch.cuda.synchronize()
t = dbg.view(148, 8).cpu().double()
used = t[:, 0] > 0
t = t[used]
t0 = t[:, 0].min()
names = ["start", "setup", "phase1 done", "panel acquired", "first full", "first acc ready", "last tile stored", "exit"]
print(f"{M}x{N}x{K} mode={sys.argv[4] if len(sys.argv) > 4 else 'round'} CTAs={int(used.sum())}")
for i, n in enumerate(names):
    col = t[:, i]
    col = col[col > 0] - t0
    if len(col):
        print(f"  {n:18s} min {col.min()/1e3:7.2f} us  median {col.median()/1e3:7.2f} us  max {col.max()/1e3:7.2f} us  (n={len(col)})")
