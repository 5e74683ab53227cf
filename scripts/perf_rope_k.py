"""RoPE-epilogue cost fit: q|k|v shape (N = 12288, M = 2048), K sweep: python scripts/perf_rope_k.py"""
import sys
import torch
sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L
from scripts.perf_epi_k import timeit  # noqa
dev = torch.device("cuda:0")
M, N, S = 2048, 12288, 2048
cs = torch.full((N,), 3e-6, device=dev)
ang = torch.outer(torch.arange(S, dtype=torch.float32), 1.0 / (10000.0 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128)))
emb = torch.cat([ang, ang], dim=-1)
cos, sin = L.rope_tables_blocked(emb.cos().bfloat16().to(dev)), L.rope_tables_blocked(emb.sin().bfloat16().to(dev))
print("K     | plain cs | rope cs | rope scalar | rope v-only(cols=0) | rope cs dup  (us per launch)")
for K in (256, 1024, 4096):
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    w = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    r = [timeit(lambda: L.w8a8_linear_q8(a, w, None, 1.0, col_scale=cs)),
         timeit(lambda: L.w8a8_linear_q8(a, w, None, 1.0, col_scale=cs, rope=(cos, sin, S, 8192))),
         timeit(lambda: L.w8a8_linear_q8(a, w, None, 3e-6, rope=(cos, sin, S, 8192))),
         timeit(lambda: L.w8a8_linear_q8(a, w, None, 3e-6, rope=(cos, sin, S, 0))),
         timeit(lambda: L.w8a8_linear_q8(a, w, None, 1.0, col_scale=cs, rope=(cos, sin, S, 8192, True)))]
    print(f"{K:5d} | " + " | ".join(f"{v:7.1f}" for v in r), flush=True)
