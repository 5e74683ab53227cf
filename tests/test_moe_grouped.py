"""Mixtral sparse-MoE block: the grouped two-launch path against the reference's per-expert loop
(HF 4.42 MixtralSparseMoeBlock.forward over Int8MixtralBlockSparseTop2MLP, models/mixtral.py:94-159).

* every expert output row of the grouped launches == the expert's own module on that token (bit-exact) when the
  loop uses the same SiLU arithmetic (asq_silu_mul_quant's SFU exp / reciprocal);
* against the loop with torch's F.silu the block output agrees within the documented SiLU tolerance (rare one-ulp
  differences of the bf16 product);
* routing helper: every slot lands in its expert's 256-row-aligned segment, padding blocks are -1.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from autosmoothquant_b200 import _lib as L
    from autosmoothquant_b200 import moe
    from autosmoothquant_b200.layers.nn.linear import W8A8BFP32OFP32Linear, W8A8BFP32OFP32LinearWithQuantScale

DEV = torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def make_experts(E, hidden, ffn, fc1, fc2, seed):
    g = torch.Generator().manual_seed(seed)
    w1, w3, w2 = [], [], []
    for e in range(E):
        def lin(i, o):
            m = torch.nn.Linear(i, o, bias=False)
            with torch.no_grad():
                m.weight.copy_(torch.randn(o, i, generator=g) * (0.04 + 0.01 * e))
            return m
        w1.append(W8A8BFP32OFP32Linear.from_float(lin(hidden, ffn), 0.03, act_quant=fc1).to(DEV))
        w3.append(W8A8BFP32OFP32Linear.from_float(lin(hidden, ffn), 0.03, act_quant=fc1).to(DEV))
        w2.append(W8A8BFP32OFP32LinearWithQuantScale.from_float(lin(ffn, hidden), 0.02 + 0.005 * e, act_quant=fc2).to(DEV))
    return w1, w3, w2


def loop_with_glue_silu(h, gate_w, w1, w3, w2, top_k):
    """Reference loop, SiLU*up evaluated by asq_silu_mul_quant (the arithmetic the epilogue shares)."""
    router_logits = F.linear(h, gate_w)
    routing = F.softmax(router_logits, dim=1, dtype=torch.float)
    routing, selected = torch.topk(routing, top_k, dim=-1)
    routing = (routing / routing.sum(dim=-1, keepdim=True)).to(h.dtype)
    out = torch.zeros_like(h)
    mask = F.one_hot(selected, num_classes=len(w1)).permute(2, 1, 0)
    for e in range(len(w1)):
        idx, top_x = torch.where(mask[e])
        if top_x.numel() == 0:
            continue
        cur = h[top_x]
        gu = torch.cat([w1[e](cur), w3[e](cur)], dim=-1).contiguous()
        _, a = L.silu_mul_quant(gu, 1.0, want_q=False, want_a=True)
        out.index_add_(0, top_x, (w2[e](a) * routing[top_x, idx, None]).to(h.dtype))
    return out


@pytest.mark.parametrize("fc1,fc2", [("per-token", "per-token"), ("per-tensor", "per-tensor")])
@pytest.mark.parametrize("T,E,hidden,ffn", [(300, 8, 512, 1024), (2048, 8, 1024, 1792), (5, 4, 256, 320)])
def test_grouped_moe_block_equals_expert_loop(fc1, fc2, T, E, hidden, ffn):
    w1, w3, w2 = make_experts(E, hidden, ffn, fc1, fc2, seed=T + E)
    g = torch.Generator().manual_seed(3)
    scale = 30.0 if fc1 == "per-tensor" else 1.0  # per-tensor fc1 rounds its (norm-folded) input directly
    h = (torch.randn(T, hidden, generator=g) * scale).to(torch.bfloat16).to(DEV)
    gate_w = (torch.randn(E, hidden, generator=g) * 0.1).to(torch.bfloat16).to(DEV)
    experts = moe.GroupedInt8Experts(w1, w3, w2)
    before = L.launch_count()
    got, logits = moe.sparse_moe_forward(h, gate_w, experts, top_k=2)
    assert L.launch_count() - before == 2  # the whole block: two kernel launches of ours
    want = loop_with_glue_silu(h, gate_w, w1, w3, w2, 2)
    torch.cuda.synchronize()
    assert got.shape == h.shape and torch.isfinite(got).all() and float(got.abs().max()) > 0
    assert torch.equal(got, want), f"{(got != want).sum().item()} of {got.numel()} elements differ from the expert loop"
    # torch's SiLU instead of the SFU one: same result up to rare one-ulp differences of the product
    ref, ref_logits = moe.sparse_moe_forward_loop(h, gate_w, w1, w3, w2, 2)
    assert torch.equal(logits, ref_logits)
    diff = (got.float() - ref.float()).abs()
    assert float((diff > 0).float().mean()) < 0.05
    assert float(diff.max()) <= 0.04 * float(ref.float().abs().max())


def test_route_tokens_layout():
    g = torch.Generator().manual_seed(0)
    E, T, k = 8, 777, 2
    sel = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(T)]).to(DEV)
    sel[:, :] = torch.where(sel == 5, torch.full_like(sel, 6), sel)  # expert 5 gets no tokens at all
    sel[:, 1] = torch.where(sel[:, 1] == sel[:, 0], (sel[:, 1] + 1) % E, sel[:, 1])
    dest, blk, m_pad = moe.route_tokens(sel, E)
    dest, blk = dest.cpu(), blk.cpu()
    flat = sel.reshape(-1).cpu()
    assert m_pad % 256 == 0 and blk.numel() == m_pad // 128
    assert dest.unique().numel() == dest.numel() and int(dest.max()) < m_pad
    assert torch.equal(blk[dest // 128].long(), flat)  # every slot sits in a block of its own expert
    counts = torch.bincount(flat, minlength=E)
    used = int(((counts + 255) // 256 * 256).sum())
    assert (blk[used // 128:] == -1).all() and (blk[:used // 128] >= 0).all()
    assert not (blk == 5).any()
    for e in range(E):  # original token order is kept inside an expert
        rows = dest[flat == e]
        assert torch.equal(rows, torch.sort(rows).values)
