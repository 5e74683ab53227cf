"""Mixtral sparse-MoE block on the grouped W8A8 kernel (BASELINE config 4).

Reference: ``Int8MixtralSparseMoeBlock`` (``autosmoothquant/models/mixtral.py:124-159``) borrows HF 4.42's
``MixtralSparseMoeBlock.forward`` — softmax(fp32) -> top-k -> renormalise -> for every expert: gather its
tokens, run ``Int8MixtralBlockSparseTop2MLP`` (``w2(act(w1 x) * w3 x)``, ``mixtral.py:94-121``), scale by the
routing weight, ``index_add_`` into the output — i.e. 3 quantized-linear launches (plus ~10 eager ones each)
PER EXPERT.  Here the routed rows are sorted by expert once, every expert's segment is padded to 256 rows, and
the whole block is TWO launches: w1|w3 of all experts with the SwiGLU product in the epilogue, then w2 of all
experts.  Rows are independent in both (per-token and per-tensor quantisation are row-local, the integer GEMM is
exact), so every expert output row equals what the expert's own module returns for that token bit for bit, and
with top-2 the final bf16 accumulation is order-independent: the block output is bit-identical to the loop.

Routing (a [T, E] matmul, softmax, top-k) and the final scatter are ordinary torch ops: not INT8 work.
No host synchronisation: buffers are sized for the worst case and unused 128-row blocks are marked -1 (the
kernel skips their tiles).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .layers.nn.linear import W8A8BFP32OFP32Linear, W8A8BFP32OFP32LinearWithQuantScale

SEGMENT_ALIGN = 256  # rows: one CTA-pair tile


class GroupedInt8Experts(nn.Module):
    """The experts of one MoE block stacked for the grouped kernel.  Built from per-expert reference-API modules
    (``w1``/``w3``: W8A8BFP32OFP32Linear, ``w2``: W8A8BFP32OFP32LinearWithQuantScale; no bias, as in Mixtral)."""

    def __init__(self, w1: Sequence[W8A8BFP32OFP32Linear], w3: Sequence[W8A8BFP32OFP32Linear],
                 w2: Sequence[W8A8BFP32OFP32LinearWithQuantScale]):
        super().__init__()
        E = len(w1)
        if not (E == len(w2) == len(w3)) or E == 0:
            raise ValueError("need the same number (>= 1) of w1 / w2 / w3 modules")
        if any(m.use_bias for m in (*w1, *w2, *w3)):
            raise NotImplementedError("grouped experts: bias is not supported (Mixtral experts have none)")
        self.num_experts = E
        self.hidden, self.ffn = w1[0].in_features, w1[0].out_features
        if self.ffn % 32:
            raise ValueError("ffn dimension must be a multiple of 32 (interleaved w1|w3 layout)")
        self.fc1_act, self.fc2_act = w1[0].act_quant, w2[0].act_quant
        dev = w1[0].weight.device
        # w1|w3 of every expert in the interleaved layout of the SwiGLU epilogue, stacked along N
        self.register_buffer("w13", torch.cat([_lib.interleave_gate_up(a.weight, b.weight) for a, b in zip(w1, w3)]).contiguous())
        self.register_buffer("w2", torch.cat([m.weight for m in w2]).contiguous())
        f32 = dict(dtype=torch.float32, device=dev)
        self.register_buffer("w1_scale", torch.tensor([float(m.dequant_scale) for m in w1], **f32))
        self.register_buffer("w3_scale", torch.tensor([float(m.dequant_scale) for m in w3], **f32))
        self.register_buffer("w2_scale", torch.tensor([float(m.dequant_scale) for m in w2], **f32))
        self.register_buffer("w2_quant_scale", torch.tensor([float(getattr(m, "quant_scale", 1.0)) for m in w2], **f32))  # per-tensor fc2 only

    @torch.no_grad()
    def forward(self, x_sorted: torch.Tensor, group_of_blk: torch.Tensor, tp_group=None, tp_world: int = 1,
                local_scales: bool = False) -> torch.Tensor:
        """x_sorted [M_pad, hidden] (rows sorted by expert, segments padded to 256) -> [M_pad, hidden].
        tp_world > 1: this module holds the Megatron shard of every expert (w1 / w3 split by column, w2 by row) and
        the result is this rank's PARTIAL sum over its ffn slice.  Per-token fc2 then needs the row absmax over the
        whole ffn dimension (SURVEY 8(e) wrinkle): the local per-token scales are max-all-reduced and w2 quantises
        with the supplied scales, which keeps the integer operands identical to the unsharded module
        (local_scales=True skips the exchange: valid, not bit-identical)."""
        mode1 = _lib.ACT_PER_TOKEN if self.fc1_act == "per-token" else _lib.ACT_ROUND
        a = _lib.w8a8_grouped_linear(x_sorted, self.w13, group_of_blk, self.w1_scale, mode1,
                                     group_dequant_scale_up=self.w3_scale, swiglu=True)
        if self.fc2_act == "per-token":
            if tp_world > 1 and not local_scales:
                import torch.distributed as dist

                _, row_scale = _lib.quantize_act(a, _lib.ACT_PER_TOKEN)
                dist.all_reduce(row_scale, op=dist.ReduceOp.MAX, group=tp_group)
                return _lib.w8a8_grouped_linear(a, self.w2, group_of_blk, self.w2_scale, _lib.ACT_ROW_SCALE_GIVEN,
                                                row_scale=row_scale)
            return _lib.w8a8_grouped_linear(a, self.w2, group_of_blk, self.w2_scale, _lib.ACT_PER_TOKEN)
        return _lib.w8a8_grouped_linear(a, self.w2, group_of_blk, self.w2_scale, _lib.ACT_SCALE,
                                        group_quant_scale=self.w2_quant_scale)


def route_tokens(selected_experts: torch.Tensor, num_experts: int):
    """selected_experts [T, top_k] -> (dest_row [T*top_k], group_of_blk int32 [M_pad/128], M_pad) for the
    expert-sorted, 256-row padded layout.  Pure device code (no host sync): M_pad is the worst case."""
    flat = selected_experts.reshape(-1)
    n = flat.numel()
    m_pad = (n + num_experts * (SEGMENT_ALIGN - 1)) // SEGMENT_ALIGN * SEGMENT_ALIGN
    # one-hot sum instead of torch.bincount: bincount sizes its output from the data (a host synchronisation, and an
    # error under CUDA-graph capture); this form has a static shape
    counts = F.one_hot(flat, num_classes=num_experts).sum(dim=0)
    padded = (counts + SEGMENT_ALIGN - 1) // SEGMENT_ALIGN * SEGMENT_ALIGN
    seg_end = torch.cumsum(padded, 0)
    seg_start = seg_end - padded
    order = torch.argsort(flat, stable=True)  # slots grouped by expert, original order inside an expert
    rank_in_expert = torch.arange(n, device=flat.device) - torch.repeat_interleave(torch.cumsum(counts, 0) - counts, counts,
                                                                                   output_size=n)
    dest_sorted = seg_start[flat[order]] + rank_in_expert
    dest_row = torch.empty_like(dest_sorted)
    dest_row[order] = dest_sorted
    blk_start = torch.arange(m_pad // 128, device=flat.device) * 128
    grp = torch.searchsorted(seg_end, blk_start, right=True)
    group_of_blk = torch.where(blk_start < seg_end[-1], grp, torch.full_like(grp, -1)).to(torch.int32)
    return dest_row, group_of_blk.contiguous(), m_pad


@torch.no_grad()
def sparse_moe_forward(hidden_states: torch.Tensor, gate_weight: torch.Tensor, experts: GroupedInt8Experts,
                       top_k: int = 2, tp_group=None, tp_world: int = 1) -> torch.Tensor:
    """HF 4.42 MixtralSparseMoeBlock.forward with the expert loop replaced by two grouped launches.
    tp_world > 1: `experts` is this rank's shard; the returned block output is a partial sum the caller
    all-reduces once (routing and the scatter are replicated on every rank)."""
    shape = hidden_states.shape
    h = hidden_states.reshape(-1, shape[-1])
    T = h.shape[0]
    router_logits = F.linear(h, gate_weight)
    routing = F.softmax(router_logits, dim=1, dtype=torch.float)
    routing, selected = torch.topk(routing, top_k, dim=-1)
    routing = (routing / routing.sum(dim=-1, keepdim=True)).to(h.dtype)
    dest_row, group_of_blk, m_pad = route_tokens(selected, experts.num_experts)
    x_sorted = torch.zeros((m_pad, h.shape[1]), dtype=h.dtype, device=h.device)
    token_of_slot = torch.arange(T, device=h.device).repeat_interleave(top_k)
    x_sorted[dest_row] = h[token_of_slot]
    y_sorted = experts(x_sorted, group_of_blk, tp_group=tp_group, tp_world=tp_world)
    weighted = y_sorted[dest_row] * routing.reshape(-1, 1)  # expert output * routing weight, in the activation dtype
    out = torch.zeros_like(h)
    out.index_add_(0, token_of_slot, weighted)
    return out.view(shape), router_logits


@torch.no_grad()
def sparse_moe_forward_loop(hidden_states: torch.Tensor, gate_weight: torch.Tensor, w1: List[nn.Module], w3: List[nn.Module],
                            w2: List[nn.Module], top_k: int = 2) -> torch.Tensor:
    """The reference's formulation (per-expert loop over the unmodified module classes); used as the parity
    baseline of the grouped path and as the 'before' of the launch-count comparison."""
    shape = hidden_states.shape
    h = hidden_states.reshape(-1, shape[-1])
    router_logits = F.linear(h, gate_weight)
    routing = F.softmax(router_logits, dim=1, dtype=torch.float)
    routing, selected = torch.topk(routing, top_k, dim=-1)
    routing = (routing / routing.sum(dim=-1, keepdim=True)).to(h.dtype)
    out = torch.zeros_like(h)
    mask = F.one_hot(selected, num_classes=len(w1)).permute(2, 1, 0)
    for e in range(len(w1)):
        idx, top_x = torch.where(mask[e])
        if top_x.numel() == 0:
            continue
        cur = h[None, top_x].reshape(-1, h.shape[1])
        cur = w2[e](F.silu(w1[e](cur)) * w3[e](cur)) * routing[top_x, idx, None]
        out.index_add_(0, top_x, cur.to(h.dtype))
    return out.view(shape), router_logits
