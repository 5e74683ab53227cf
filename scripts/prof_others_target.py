"""ncu target for the kernels BESIDES the per-tensor Llama layer (profiles/r01e): one launch each, inside a
cudaProfilerStart/Stop range, after a warm launch outside it.

    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/<tag>/others python scripts/prof_others_target.py

 #0  FP8-e4m3 per-token module forward (BASELINE config 5 granularity)      2048 x 8192 x 8192   kind::f8f6f4, fused prologue
 #1  INT8 per-token module forward (config 3: Llama-2-13B down_proj)          8192 x 5120 x 13824  phase 1 with row scales
 #2  grouped MoE w1|w3 + SwiGLU, per-token (config 4: Mixtral-8x7B, top-2)    4096 routed rows (+ padding) x 2*14336 x 4096, 8 experts
 #3  grouped MoE w2, per-token                                                 the same rows x 4096 x 14336
 #4  batched INT8 GEMM (csrc/kernels/bmm.cu family): 32 x (2048 x 2048 x 128)  s8 x s8 -> f32
"""
import sys

import torch

sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L, moe  # noqa: E402
from autosmoothquant_b200.layers.nn.linear import W8A8BFP32OFP32Linear, W8A8BFP32OFP32LinearWithQuantScale  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)


def randn(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)


# 0: FP8 per-token
x_f8 = randn(2048, 8192)
w_f8 = (torch.randn(8192, 8192, device=dev, generator=g) * 20).clamp(-448, 448).to(torch.float8_e4m3fn)
# 1: INT8 per-token, 13B down_proj
x_pt = randn(8192, 13824)
w_pt = torch.randint(-127, 128, (5120, 13824), dtype=torch.int8, device=dev, generator=g)
# 2, 3: Mixtral experts
E, H, FFN, T = 8, 4096, 14336, 2048


def mk(cls, i, o):
    m = cls(i, o, False, "per-token")
    m.weight = torch.randint(-127, 128, (o, i), dtype=torch.int8, device=dev, generator=g)
    m.dequant_scale = torch.tensor(3e-4)
    return m.to(dev)


experts = moe.GroupedInt8Experts([mk(W8A8BFP32OFP32Linear, H, FFN) for _ in range(E)],
                                 [mk(W8A8BFP32OFP32Linear, H, FFN) for _ in range(E)],
                                 [mk(W8A8BFP32OFP32LinearWithQuantScale, FFN, H) for _ in range(E)])
torch.cuda.empty_cache()
gate = randn(E, H, scale=0.05)
h = randn(T, H)
sel = torch.topk(torch.softmax(torch.nn.functional.linear(h, gate).float(), 1), 2)[1]
dest, blk, m_pad = moe.route_tokens(sel, E)
xs = torch.zeros(m_pad, H, dtype=torch.bfloat16, device=dev)
xs[dest] = h[torch.arange(T, device=dev).repeat_interleave(2)]
# 4: batched GEMM
a_b = torch.randint(-128, 128, (32, 2048, 128), dtype=torch.int8, device=dev, generator=g)
b_b = torch.randint(-128, 128, (32, 2048, 128), dtype=torch.int8, device=dev, generator=g)


def all_launches():
    L.fp8_linear(x_f8, w_f8, None, L.ACT_PER_TOKEN, 1.0, 0.01)
    L.w8a8_linear(x_pt, w_pt, None, L.ACT_PER_TOKEN, 1.0, 3e-4)
    experts(xs, blk)  # two launches
    L.i8bmm(a_b, b_b, torch.float32, 0.001)


all_launches()
torch.cuda.synchronize()
n0 = L.launch_count()
torch.cuda.profiler.start()
all_launches()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"profiled {L.launch_count() - n0} launches; m_pad {m_pad}, real routed rows {T * 2}")
