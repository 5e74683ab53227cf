"""Mixtral-8x7B sparse-MoE block (BASELINE config 4: experts per-token, TP=8 shard: ffn 14336/8 = 1792 per rank,
hidden 4096, 8 experts, top-2, 2048 tokens): grouped two-launch path vs the per-expert loop. python scripts/perf_moe.py"""
import sys
import torch
sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L, moe
from autosmoothquant_b200.layers.nn.linear import W8A8BFP32OFP32Linear, W8A8BFP32OFP32LinearWithQuantScale
dev = torch.device("cuda:0")

def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for ffn in (1792, 14336):
    E, H, T = 8, 4096, 2048
    def mk(cls, i, o, act):
        m = cls(i, o, False, act)
        m.weight = torch.randint(-127, 128, (o, i), dtype=torch.int8)
        m.dequant_scale = torch.tensor(3e-4)
        return m.to(dev)
    w1 = [mk(W8A8BFP32OFP32Linear, H, ffn, "per-token") for _ in range(E)]
    w3 = [mk(W8A8BFP32OFP32Linear, H, ffn, "per-token") for _ in range(E)]
    w2 = [mk(W8A8BFP32OFP32LinearWithQuantScale, ffn, H, "per-token") for _ in range(E)]
    h = torch.randn(T, H, device=dev).to(torch.bfloat16)
    gate = (torch.randn(E, H, device=dev) * 0.05).to(torch.bfloat16)
    experts = moe.GroupedInt8Experts(w1, w3, w2)
    t_loop = timeit(lambda: moe.sparse_moe_forward_loop(h, gate, w1, w3, w2, 2))
    t_grp = timeit(lambda: moe.sparse_moe_forward(h, gate, experts, 2))
    sel = torch.topk(torch.softmax(torch.nn.functional.linear(h, gate).float(), 1), 2)[1]
    dest, blk, m_pad = moe.route_tokens(sel, E)
    xs = torch.zeros(m_pad, H, dtype=torch.bfloat16, device=dev)
    t_kern = timeit(lambda: experts(xs, blk))
    ops = 2.0 * T * 2 * 3 * H * ffn
    print(f"ffn/rank {ffn}: expert loop {t_loop:8.1f} us | grouped block {t_grp:8.1f} us | the 2 grouped launches alone {t_kern:8.1f} us "
          f"({ops / t_kern / 1e6:.0f} TOPS on the {T * 2} real rows, M_pad {m_pad})", flush=True)
    del w1, w2, w3, experts
    torch.cuda.empty_cache()
