import sys, torch
sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L
dev = torch.device("cuda:0")
def timeit(fn, iters=50, warmup=5):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
M, H, I = 2048, 4096, 11008
bufs = [torch.randn(M, 2 * I, device=dev).to(torch.bfloat16) for _ in range(3)]
i = [0]
def silu():
    i[0] += 1; return L.silu_mul_quant(bufs[i[0] % 3], 0.06)
print(f"silu_mul_quant {timeit(silu):.1f} us  ({(M*2*I*2 + M*I)/1e6:.0f} MB)")
xs = [torch.randn(M, H, device=dev).to(torch.bfloat16) for _ in range(8)]
w = torch.ones(H, device=dev, dtype=torch.bfloat16)
def norm():
    i[0] += 1; return L.add_rmsnorm_quant(xs[i[0] % 8], xs[(i[0] + 1) % 8], w, 1e-5)
print(f"add_rmsnorm_quant {timeit(norm):.1f} us  ({M*H*7/1e6:.0f} MB)")
qkv = [torch.randn(M, 3 * H, device=dev).to(torch.bfloat16) for _ in range(3)]
cos = torch.randn(M, 128, device=dev).to(torch.bfloat16); sin = cos.clone()
def rope():
    i[0] += 1; L.rope_inplace(qkv[i[0] % 3], cos, sin, M, 64, 128)
print(f"rope_inplace {timeit(rope):.1f} us  ({M*2*H*2*2/1e6:.0f} MB)")
