"""Tiny decoder-only language models with the MODULE NAMES of the four families the reference quantizes.

TEST INFRASTRUCTURE (shared by oracle/gen_golden_calib.py and tests/): the reference's offline pipeline addresses
modules by name (``model.layers.3.self_attn.q_proj``, ``model.decoder.layers.0.fc1``,
``model.layers.1.block_sparse_moe.experts.5.w2`` ...; autosmoothquant/quantize/calibration.py:90-183,
quantize/smooth.py:42-93) and hooks every ``nn.Linear``.  These models reproduce that naming and data flow — a norm
feeding q/k/v (or a packed W_pack), an attention output projection, a norm feeding the MLP / the router and experts —
at a size that runs in milliseconds on CPU.  They are not the HF implementations (no RoPE, no causal mask): the
calibration code under test only sees Linear inputs and outputs.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn

ARCHITECTURE = {"llama": "LlamaForCausalLM", "baichuan": "BaichuanForCausalLM", "mixtral": "MixtralForCausalLM",
                "transformers": "OPTForCausalLM"}


class RMSNorm(nn.Module):
    def __init__(self, n, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.variance_epsilon = eps

    def forward(self, x):
        return self.weight * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.variance_epsilon))


def _attend(q, k, v, heads):
    b, s, h = q.shape
    split = lambda t: t.view(b, s, heads, h // heads).transpose(1, 2)
    p = torch.softmax(split(q) @ split(k).transpose(-1, -2) / (h // heads) ** 0.5, dim=-1)
    return (p @ split(v)).transpose(1, 2).reshape(b, s, h)


class Attention(nn.Module):
    def __init__(self, hidden, heads, packed=False, out_name="o_proj", bias=False):
        super().__init__()
        self.heads, self.packed, self.out_name = heads, packed, out_name
        if packed:
            self.W_pack = nn.Linear(hidden, 3 * hidden, bias=False)
        else:
            self.q_proj = nn.Linear(hidden, hidden, bias=bias)
            self.k_proj = nn.Linear(hidden, hidden, bias=bias)
            self.v_proj = nn.Linear(hidden, hidden, bias=bias)
        setattr(self, out_name, nn.Linear(hidden, hidden, bias=bias))

    def forward(self, x):
        if self.packed:
            q, k, v = self.W_pack(x).chunk(3, dim=-1)
        else:
            q, k, v = self.q_proj(x), self.k_proj(x), self.v_proj(x)
        return getattr(self, self.out_name)(_attend(q, k, v, self.heads))


class GatedMLP(nn.Module):
    def __init__(self, hidden, inter):
        super().__init__()
        self.gate_proj = nn.Linear(hidden, inter, bias=False)
        self.up_proj = nn.Linear(hidden, inter, bias=False)
        self.down_proj = nn.Linear(inter, hidden, bias=False)

    def forward(self, x):
        return self.down_proj(nn.functional.silu(self.gate_proj(x)) * self.up_proj(x))


class Expert(nn.Module):
    def __init__(self, hidden, inter):
        super().__init__()
        self.w1 = nn.Linear(hidden, inter, bias=False)
        self.w2 = nn.Linear(inter, hidden, bias=False)
        self.w3 = nn.Linear(hidden, inter, bias=False)

    def forward(self, x):
        return self.w2(nn.functional.silu(self.w1(x)) * self.w3(x))


class SparseMoeBlock(nn.Module):
    """Router + experts with HF Mixtral's routing rule; an expert runs only on the tokens routed to it, and ``top_k``
    is a mutable attribute (the reference raises it to the expert count while calibrating, calibration.py:23-42)."""

    def __init__(self, hidden, inter, experts, top_k):
        super().__init__()
        self.top_k = top_k
        self.gate = nn.Linear(hidden, experts, bias=False)
        self.experts = nn.ModuleList([Expert(hidden, inter) for _ in range(experts)])

    def forward(self, x):
        shape = x.shape
        x = x.reshape(-1, shape[-1])
        probs = torch.softmax(self.gate(x).float(), dim=-1)
        weight, chosen = torch.topk(probs, self.top_k, dim=-1)
        weight = (weight / weight.sum(-1, keepdim=True)).to(x.dtype)
        out = torch.zeros_like(x)
        for e, expert in enumerate(self.experts):
            rows, slot = torch.where(chosen == e)
            if rows.numel():
                out.index_add_(0, rows, expert(x[rows]) * weight[rows, slot, None])
        return out.view(shape)


class LlamaLikeLayer(nn.Module):
    def __init__(self, kind, hidden, inter, heads, experts, top_k):
        super().__init__()
        self.input_layernorm = RMSNorm(hidden)
        self.self_attn = Attention(hidden, heads, packed=(kind == "baichuan"))
        self.post_attention_layernorm = RMSNorm(hidden)
        if kind == "mixtral":
            self.block_sparse_moe = SparseMoeBlock(hidden, inter, experts, top_k)
        else:
            self.mlp = GatedMLP(hidden, inter)

    def forward(self, x):
        x = x + self.self_attn(self.input_layernorm(x))
        ffn = self.block_sparse_moe if hasattr(self, "block_sparse_moe") else self.mlp
        return x + ffn(self.post_attention_layernorm(x))


class OPTLikeLayer(nn.Module):
    def __init__(self, hidden, inter, heads):
        super().__init__()
        self.self_attn_layer_norm = nn.LayerNorm(hidden)
        self.self_attn = Attention(hidden, heads, out_name="out_proj", bias=True)
        self.final_layer_norm = nn.LayerNorm(hidden)
        self.fc1 = nn.Linear(hidden, inter)
        self.fc2 = nn.Linear(inter, hidden)

    def forward(self, x):
        x = x + self.self_attn(self.self_attn_layer_norm(x))
        return x + self.fc2(torch.relu(self.fc1(self.final_layer_norm(x))))


class _Stack(nn.Module):
    def __init__(self, layers, vocab, hidden):
        super().__init__()
        self.embed_tokens = nn.Embedding(vocab, hidden)
        self.layers = nn.ModuleList(layers)

    def forward(self, ids):
        x = self.embed_tokens(ids)
        for layer in self.layers:
            x = layer(x)
        return x


class _OPTModel(nn.Module):  # OPT nests its stack one level deeper: model.decoder.layers
    def __init__(self, stack):
        super().__init__()
        self.decoder = stack

    def forward(self, ids):
        return self.decoder(ids)


class TinyLM(nn.Module):
    def __init__(self, kind, vocab=97, hidden=32, inter=48, heads=4, layers=2, experts=4, top_k=2, seed=0):
        super().__init__()
        if kind not in ARCHITECTURE:
            raise ValueError(kind)
        torch.manual_seed(seed)
        if kind == "transformers":
            self.model = _OPTModel(_Stack([OPTLikeLayer(hidden, inter, heads) for _ in range(layers)], vocab, hidden))
        else:
            self.model = _Stack([LlamaLikeLayer(kind, hidden, inter, heads, experts, top_k) for _ in range(layers)], vocab, hidden)
        self.lm_head = nn.Linear(hidden, vocab, bias=False)
        self.config = SimpleNamespace(architectures=[ARCHITECTURE[kind]], num_hidden_layers=layers, hidden_size=hidden,
                                      num_local_experts=experts, num_experts_per_tok=top_k, pretraining_tp=1)
        with torch.no_grad():  # a few outlier channels after every norm: the regime SmoothQuant is for
            for m in self.modules():
                if isinstance(m, (RMSNorm, nn.LayerNorm)):
                    m.weight[::7] *= 6.0

    def forward(self, ids):
        return self.lm_head(self.model(ids))


def calibration_batches(n=6, seq=24, vocab=97, seed=1):
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(0, vocab, (1, seq), generator=g) for _ in range(n)]
