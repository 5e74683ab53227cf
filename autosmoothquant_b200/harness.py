"""Minimal Llama-style decoder stack used to measure the W8A8 linear path at model level.

This is measurement plumbing, not a model zoo: the reference's model classes subclass HF
transformers 4.42 internals that no longer exist in the installed transformers, so the
benchmark needs its own thin stack.  Every projection is one of the reference-API modules from
``autosmoothquant_b200.layers.nn.linear`` chosen from a ``quant_config`` dict exactly as the
reference's ``autosmoothquant/models/llama.py:74-106,185-214`` does, with the norm-weight folding
of ``llama.py:27-37,326-339``; everything that is not a quantized linear (embedding, RMSNorm,
RoPE, attention, SiLU, lm_head) is ordinary torch and is reported separately from the INT8 work.

Weights are synthetic (seeded normal, true shapes), converted with the modules' own
``from_float`` (the reference's absmax weight quantiser).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

from .layers.nn.linear import (
    FP8LinearDynamic,
    W8A8BFP32OFP32Linear,
    W8A8BFP32OFP32LinearWithQuantScale,
    W8A8BFP32OFP32QKVLinear,
)


@dataclass(frozen=True)
class DecoderConfig:
    name: str
    hidden: int
    intermediate: int
    layers: int
    heads: int
    kv_heads: int
    vocab: int = 32000
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    experts: int = 0  # > 0: Mixtral-style sparse-MoE MLP (every expert: w1 | w3 -> SwiGLU -> w2), top_k routed
    top_k: int = 2

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    def linear_macs_per_token_per_layer(self) -> int:
        """MACs of the QUANTIZED linears per token (the router of an MoE block is a plain 16-bit matmul)."""
        kv = self.kv_heads * self.head_dim
        mlp = 3 * self.hidden * self.intermediate * (self.top_k if self.experts else 1)
        return 2 * self.hidden * self.hidden + 2 * self.hidden * kv + mlp


LLAMA2_7B = DecoderConfig("llama-2-7b", 4096, 11008, 32, 32, 32)
LLAMA2_13B = DecoderConfig("llama-2-13b", 5120, 13824, 40, 40, 40)
LLAMA2_70B = DecoderConfig("llama-2-70b", 8192, 28672, 80, 64, 8)
MIXTRAL_8X7B = DecoderConfig("mixtral-8x7b", 4096, 14336, 32, 32, 8, rope_theta=1e6, experts=8, top_k=2)
TINY = DecoderConfig("tiny", 256, 512, 2, 4, 4, vocab=512)
TINY_GQA = DecoderConfig("tiny-gqa", 512, 1024, 2, 4, 2, vocab=512)  # head_dim 128: exercises the RoPE epilogue
TINY_MOE = DecoderConfig("tiny-moe", 512, 1024, 2, 4, 2, vocab=512, experts=4, top_k=2)
CONFIGS = {c.name: c for c in (LLAMA2_7B, LLAMA2_13B, LLAMA2_70B, MIXTRAL_8X7B, TINY, TINY_GQA, TINY_MOE)}

DEFAULT_QUANT_CONFIG = {"qkv": "per-tensor", "out": "per-tensor", "fc1": "per-tensor", "fc2": "per-tensor",
                        "type": "int8"}

# synthetic calibration results (activation absmax / 127) for seeded N(0, 0.02) weights
SYNTH_INPUT_SCALES = {"attn_input": 4.5 / 127, "out_input": 6.0 / 127, "gate_input": 4.5 / 127, "down_input": 8.0 / 127}


def normalise_quant_config(cfg: Dict[str, str]) -> Dict[str, str]:
    """quant_config.json semantics (reference README.md:25-41, smoothquant_model.py:62-70):
    'fp8' is shorthand for 'fp8_e4m3'; fp8 needs an activation_scheme (default dynamic)."""
    out = dict(DEFAULT_QUANT_CONFIG)
    out.update(cfg or {})
    if out["type"] == "fp8":
        out["type"] = "fp8_e4m3"
    if out["type"] not in ("int8", "fp8_e4m3"):
        raise ValueError(f"unsupported quant type {out['type']!r}")
    if out["type"] == "fp8_e4m3":
        out.setdefault("activation_scheme", "dynamic")
    for key in ("qkv", "out", "fc1", "fc2"):
        if out[key] not in ("per-tensor", "per-token"):
            raise ValueError(f"{key} must be per-tensor or per-token")
    return out


def _rand_linear(in_f: int, out_f: int, gen: torch.Generator, device, std: float = 0.02) -> nn.Linear:
    lin = nn.Linear(in_f, out_f, bias=False, device="meta", dtype=torch.float32)  # no default init pass over the weight
    lin.weight = nn.Parameter(torch.empty(out_f, in_f, dtype=torch.float32, device=device).normal_(0.0, std, generator=gen),
                              requires_grad=False)
    return lin


def _make_proj(kind: str, in_f: int, out_f: int, qcfg: Dict[str, str], input_scale: float, gen, device, tp=None) -> nn.Module:
    """kind in {'qkv','out','fc1','fc2'} -> module class as llama.py:74-106,185-214 picks it.  tp = (rank, world):
    the full-size seeded weight is quantised exactly as on one GPU (the weight scale is per-tensor, so every shard
    shares it) and only this rank's column (qkv / fc1) or row (out / fc2) shard is kept."""
    lin = _rand_linear(in_f, out_f, gen, device)
    gran = qcfg[kind]
    if qcfg["type"] == "fp8_e4m3":
        mod = FP8LinearDynamic.from_float(lin, 1.0, save_device=device, act_quant="per-token", reference_compat=False)
    elif kind in ("qkv", "fc1"):
        mod = W8A8BFP32OFP32Linear.from_float(lin, input_scale, save_device=device, act_quant=gran)
    else:
        mod = W8A8BFP32OFP32LinearWithQuantScale.from_float(lin, input_scale, save_device=device, act_quant=gran)
    del lin
    if tp is not None and tp[1] > 1:
        from . import tp as _tp

        mod = (_tp.shard_column if kind in ("qkv", "fc1") else _tp.shard_row)(mod, tp[0], tp[1])
    return mod.to(device) if qcfg["type"] != "fp8_e4m3" else mod._apply(lambda t: t.to(device))


def _fuse_columns(mods, device) -> W8A8BFP32OFP32QKVLinear:
    """Horizontal fusion of projections that share their input (q/k/v, gate/up) into ONE launch with the
    reference's fused-W_pack module (linear.py:132-245): weights concatenated along N, every block keeps its
    own dequant scale, so the result is bit-identical to running the projections separately while the
    activation is quantised once instead of once per projection."""
    assert 2 <= len(mods) <= 3 and all(isinstance(m, W8A8BFP32OFP32Linear) for m in mods)
    sizes = [m.out_features for m in mods] + [0] * (3 - len(mods))
    first = mods[0]
    fused = W8A8BFP32OFP32QKVLinear(sizes, first.in_features, sum(sizes), first.use_bias, first.act_quant)
    fused.weight = torch.cat([m.weight for m in mods], dim=0).contiguous()
    scales = [m.dequant_scale for m in mods] + [torch.tensor(1.0)] * (3 - len(mods))
    for name, sc in zip(fused._scale_names, scales):
        setattr(fused, name, sc.detach().clone().to(torch.float32).cpu())
    if first.use_bias:
        fused.bias = torch.cat([m.bias for m in mods]).contiguous()
    return fused.to(device)


class FP8FusedColumnsLinear(nn.Module):
    """FP8LinearDynamic per-token projections that share their input (q|k|v, gate|up) as ONE launch: e4m3 weights
    concatenated along N, every block keeps its own per-tensor weight scale as a per-output-column vector
    (`asq_fp8_linear_cs`) — the FP8 counterpart of the reference's fused-W_pack INT8 module (linear.py:132-245).  The
    activation is quantised once instead of once per projection; every output element goes through the same arithmetic
    as in the separate launches, so the result is bit-identical to running the projections one by one."""

    def __init__(self, mods):
        super().__init__()
        if not mods or any(not isinstance(m, FP8LinearDynamic) or m.act_quant != "per-token" for m in mods):
            raise ValueError("FP8FusedColumnsLinear fuses per-token FP8LinearDynamic modules")
        if len({m.in_features for m in mods}) != 1 or len({m.use_bias for m in mods}) != 1:
            raise ValueError("fused projections must share in_features and all have (or all lack) a bias")
        self.in_features = mods[0].in_features
        self.sizes = [m.out_features for m in mods]
        self.out_features = sum(self.sizes)
        self.use_bias = mods[0].use_bias
        dev = mods[0].weight.device
        self.register_buffer("weight", torch.cat([m.weight.view(torch.uint8) for m in mods], dim=0).contiguous().view(torch.float8_e4m3fn))
        self.register_buffer("col_scale", torch.cat([torch.full((m.out_features,), float(m.weight_scale), dtype=torch.float32)
                                                     for m in mods]).to(dev))
        if self.use_bias:
            self.register_buffer("bias", torch.cat([m.bias.to(torch.float32) for m in mods]).contiguous())

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import _lib

        x2 = x.reshape(-1, self.in_features)
        y = _lib.fp8_linear(x2, self.weight, self.bias if self.use_bias else None, _lib.ACT_PER_TOKEN, col_scale=self.col_scale)
        return y.view(*x.shape[:-1], self.out_features)


def rms_norm_hf(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """HF LlamaRMSNorm.forward (what the reference's QuantizedLlamaRMSNorm inherits, models/llama.py:27-37):
    normalise in fp32, round to the activation dtype, then multiply by the (folded) weight."""
    xf = x.float()
    n = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype)
    return weight * n


def _rope_tables(seq: int, head_dim: int, theta: float, device, dtype):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=device, dtype=torch.float32) / head_dim))
    ang = torch.outer(torch.arange(seq, device=device, dtype=torch.float32), inv)
    emb = torch.cat((ang, ang), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    half = x.shape[-1] // 2
    rot = torch.cat((-x[..., half:], x[..., :half]), dim=-1)
    return x * cos + rot * sin


class QuantDecoderLayer(nn.Module):
    """One decoder block.  cfg.experts == 0: Llama (gate / up / down); cfg.experts > 0: Mixtral — the MLP is the
    sparse-MoE block of models/mixtral.py:94-159 (router + experts w1 / w3 / w2, quant_config fc1 / fc2).
    tp = (rank, world): every projection keeps only this rank's Megatron shard (see _make_proj)."""

    def __init__(self, cfg: DecoderConfig, qcfg: Dict[str, str], gen, device, dtype, fuse_projections: bool = False, tp=None):
        super().__init__()
        self.cfg = cfg
        self.fused = False
        self.moe = None
        h, kv, inter = cfg.hidden, cfg.kv_heads * cfg.head_dim, cfg.intermediate
        s = SYNTH_INPUT_SCALES
        int8 = qcfg["type"] == "int8"
        self.q_proj = _make_proj("qkv", h, h, qcfg, s["attn_input"], gen, device, tp)
        self.k_proj = _make_proj("qkv", h, kv, qcfg, s["attn_input"], gen, device, tp)
        self.v_proj = _make_proj("qkv", h, kv, qcfg, s["attn_input"], gen, device, tp)
        self.o_proj = _make_proj("out", h, h, qcfg, s["out_input"], gen, device, tp)
        if cfg.experts:
            if not int8:
                raise NotImplementedError("the sparse-MoE stack is built for the INT8 modules (BASELINE config 4)")
            from .moe import GroupedInt8Experts

            w1, w3, w2 = [], [], []
            for _ in range(cfg.experts):  # experts[i].w1 / w3 / w2 as models/mixtral.py:99-101 builds them
                w1.append(_make_proj("fc1", h, inter, qcfg, s["gate_input"], gen, device, tp))
                w3.append(_make_proj("fc1", h, inter, qcfg, s["gate_input"], gen, device, tp))
                w2.append(_make_proj("fc2", inter, h, qcfg, s["down_input"], gen, device, tp))
            self.moe = GroupedInt8Experts(w1, w3, w2)  # stacked copies: the per-expert modules are dropped
            del w1, w3, w2
            self.register_buffer("router_weight", torch.empty(cfg.experts, h, dtype=dtype, device=device))
            with torch.no_grad():
                self.router_weight.normal_(0.0, 0.02, generator=gen)
        else:
            self.gate_proj = _make_proj("fc1", h, inter, qcfg, s["gate_input"], gen, device, tp)
            self.up_proj = _make_proj("fc1", h, inter, qcfg, s["gate_input"], gen, device, tp)
            self.down_proj = _make_proj("fc2", inter, h, qcfg, s["down_input"], gen, device, tp)
        ln1 = torch.ones(h, dtype=torch.float32, device=device)
        ln2 = torch.ones(h, dtype=torch.float32, device=device)
        # fold 1/input_scale into the norm weight when the consumer is per-tensor (llama.py:326-339, mixtral.py:22-30)
        if int8 and qcfg["qkv"] == "per-tensor":
            ln1 = ln1 / s["attn_input"]
        if int8 and qcfg["fc1"] == "per-tensor":
            ln2 = ln2 / s["gate_input"]
        self.register_buffer("input_layernorm_weight", ln1.to(dtype))
        self.register_buffer("post_attention_layernorm_weight", ln2.to(dtype))
        if fuse_projections and int8:
            self.qkv_sizes = [self.q_proj.out_features, self.k_proj.out_features, self.v_proj.out_features]
            self.qkv_proj = _fuse_columns([self.q_proj, self.k_proj, self.v_proj], device)
            del self.q_proj, self.k_proj, self.v_proj
            if self.moe is None:
                self.gate_up_proj = _fuse_columns([self.gate_proj, self.up_proj], device)
                del self.gate_proj, self.up_proj
            self.fused = True
        elif fuse_projections and self.moe is None and os.environ.get("ASQ_FP8_FUSE", "1") != "0":
            # FP8 per-token (BASELINE config 5): 7 -> 4 launches per layer, the [M, hidden] input quantised twice instead
            # of five times; under tensor parallelism this matters most for the narrow k / v shards (N = 128 at 70B / TP8)
            self.qkv_sizes = [self.q_proj.out_features, self.k_proj.out_features, self.v_proj.out_features]
            self.qkv_proj = FP8FusedColumnsLinear([self.q_proj, self.k_proj, self.v_proj])
            self.gate_up_proj = FP8FusedColumnsLinear([self.gate_proj, self.up_proj])
            del self.q_proj, self.k_proj, self.v_proj, self.gate_proj, self.up_proj
            self.fused = True

    def enable_swiglu_epilogue(self) -> None:
        """Re-lay the fused gate|up weight out for the SwiGLU epilogue (32-row blocks of gate and up interleaved,
        `_lib.interleave_gate_up`); the glue forward then needs no gate|up tensor and no SiLU kernel."""
        assert self.fused and self.moe is None and self.gate_up_proj.weight.shape[0] % 64 == 0
        from . import _lib

        mod = self.gate_up_proj
        I = mod.weight.shape[0] // 2
        w_il = _lib.interleave_gate_up(mod.weight[:I], mod.weight[I:])
        b_il = _lib.interleave_gate_up(mod.bias[:I], mod.bias[I:]) if mod.use_bias else None
        gate_scale, up_scale = (float(getattr(mod, n).item()) for n in mod._scale_names[:2])
        self.gate_up_il = (w_il, b_il, gate_scale, up_scale)

    def linears(self):
        attn = [self.qkv_proj] if self.fused else [self.q_proj, self.k_proj, self.v_proj]
        if self.moe is not None:
            return attn + [self.o_proj, self.moe]
        if self.fused:
            return attn + [self.o_proj, self.gate_up_proj, self.down_proj]
        return attn + [self.o_proj, self.gate_proj, self.up_proj, self.down_proj]

    def _moe_forward(self, h2: torch.Tensor) -> torch.Tensor:
        """Sparse-MoE MLP on [T, hidden]; under tensor parallelism every expert is sharded (w1 / w3 by column, w2 by
        row), routing is replicated, and the weighted expert sum — a partial sum over the ffn shards — is completed
        by ONE all-reduce per block."""
        from .moe import sparse_moe_forward

        world = getattr(self, "tp_world", 1)
        out, _ = sparse_moe_forward(h2, self.router_weight, self.moe, top_k=self.cfg.top_k,
                                    tp_group=getattr(self, "tp_group", None) if world > 1 else None, tp_world=world)
        if world > 1:
            self.tp_allreduce(out)
        return out

    def tp_allreduce(self, t: torch.Tensor) -> None:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=getattr(self, "tp_group", None))

    def forward(self, x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
        cfg = self.cfg
        B, S, H = x.shape
        h = rms_norm_hf(x, self.input_layernorm_weight, cfg.rms_eps)
        # head counts are inferred from the projection width so tensor-parallel shards (heads / p) work too
        if self.fused:
            q, k, v = self.qkv_proj(h).split(self.qkv_sizes, dim=-1)
        else:
            q, k, v = self.q_proj(h), self.k_proj(h), self.v_proj(h)
        q = q.view(B, S, -1, cfg.head_dim).transpose(1, 2)
        k = k.view(B, S, -1, cfg.head_dim).transpose(1, 2)
        v = v.view(B, S, -1, cfg.head_dim).transpose(1, 2)
        q, k = _apply_rope(q, cos, sin), _apply_rope(k, cos, sin)
        attn = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=k.shape[1] != q.shape[1])
        x = x + self.o_proj(attn.transpose(1, 2).reshape(B, S, -1))
        h = rms_norm_hf(x, self.post_attention_layernorm_weight, cfg.rms_eps)
        if self.moe is not None:
            return x + self._moe_forward(h.reshape(B * S, H)).view(B, S, H)
        if self.fused:
            gate, up = self.gate_up_proj(h).chunk(2, dim=-1)
        else:
            gate, up = self.gate_proj(h), self.up_proj(h)
        x = x + self.down_proj(F.silu(gate) * up)
        return x


class RopeTables:
    """cos / sin tables of one sequence length plus the blocked copies the RoPE epilogue reads.  Built once where the
    tables are created and carried by the model (no address-keyed cache: a recycled allocation cannot alias them)."""

    def __init__(self, seq: int, head_dim: int, theta: float, device, dtype):
        from . import _lib

        self.seq = seq
        self.cos, self.sin = _rope_tables(seq, head_dim, theta, device, dtype)
        half = head_dim // 2
        self.halves_equal = bool(torch.equal(self.cos[:, :half], self.cos[:, half:]) and torch.equal(self.sin[:, :half], self.sin[:, half:]))
        self.cos_blocked = _lib.rope_tables_blocked(self.cos) if head_dim % 8 == 0 and self.cos.is_cuda else None
        self.sin_blocked = _lib.rope_tables_blocked(self.sin) if head_dim % 8 == 0 and self.sin.is_cuda else None

    def epilogue_arg(self, rope_cols: int):
        return (self.cos_blocked, self.sin_blocked, self.seq, rope_cols, self.halves_equal)


def _row_parallel(layer, name):
    """(module holding weight / scales, tensor-parallel wrapper or None) of o_proj / down_proj."""
    mod = getattr(layer, name)
    return (mod.shard, mod) if getattr(layer, "tp_world", 1) > 1 else (mod, None)


def _attention_half_glue(layer: "QuantDecoderLayer", x2: torch.Tensor, delta: Optional[torch.Tensor], B: int, S: int,
                         rope: RopeTables):
    """First half of a decoder block with the producer-side fusions of asq_glue.cu (per-tensor qkv, fused q|k|v):
    add+RMSNorm emits the int8 input of q|k|v, RoPE runs in that GEMM's epilogue, o_proj (fed by the attention
    library kernel) quantises inside its own launch and adds the residual in its epilogue.
    x2: residual stream [M,H]; delta: the previous block's output not yet added.  Returns (x2, delta)."""
    from . import _lib

    cfg = layer.cfg
    hd = cfg.head_dim
    tp_world = getattr(layer, "tp_world", 1)
    # residual add inside the o_proj / down_proj epilogues (the residual tile is TMA-loaded into the staging buffer):
    # bit-identical, the norm kernels then read one tensor and write no copy of the stream; 12.86 vs 13.02-13.10 ms
    fuse_res = tp_world == 1 and os.environ.get("ASQ_RESIDUAL_EPILOGUE", "1") != "0"
    # RMSNorm as the prologue of the q|k|v / gate|up launches (needs the residual already added: delta is None).
    # Bit-identical and 64 launches fewer per step, but measured SLOWER (13.9 vs 13.0 ms): the prologue runs while
    # the tensor cores of that launch idle (~8 us per launch), the stand-alone norm kernel costs ~5 us -> opt-in
    norm_prologue = fuse_res and os.environ.get("ASQ_NORM_PROLOGUE", "0") == "1"
    qkv_mod = layer.qkv_proj
    nq, nk, nv = (n // hd for n in layer.qkv_sizes)
    rope_in_epilogue = hd == 128 and os.environ.get("ASQ_ROPE_EPILOGUE", "1") != "0"
    rope_arg = rope.epilogue_arg((nq + nk) * hd) if rope_in_epilogue else None
    if norm_prologue and delta is None:
        qkv = _lib.w8a8_rmsnorm_linear(x2, layer.input_layernorm_weight, cfg.rms_eps, qkv_mod.weight,
                                       qkv_mod.bias if qkv_mod.use_bias else None, 1.0,
                                       col_scale=qkv_mod._col_scale(x2.device), rope=rope_arg)
    else:
        x2, _, q8 = _lib.add_rmsnorm_quant(x2, delta, layer.input_layernorm_weight, cfg.rms_eps)
        qkv = _lib.w8a8_linear_q8(q8, qkv_mod.weight, qkv_mod.bias if qkv_mod.use_bias else None, 1.0,
                                  col_scale=qkv_mod._col_scale(x2.device), out_dtype=x2.dtype, rope=rope_arg)
    if not rope_in_epilogue:
        _lib.rope_inplace(qkv, rope.cos, rope.sin, S, nq + nk, hd)
    q, k, v = qkv.split(layer.qkv_sizes, dim=-1)
    q = q.view(B, S, nq, hd).transpose(1, 2)
    k = k.view(B, S, nk, hd).transpose(1, 2)
    v = v.view(B, S, nv, hd).transpose(1, 2)
    attn = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=nk != nq)
    attn2 = attn.transpose(1, 2).reshape(B * S, nq * hd)
    om, o_tp = _row_parallel(layer, "o_proj")
    if o_tp is not None:
        # row-parallel wrapper: quantise (global row scales when per-token), GEMM, all-reduce (tp.RowParallelLinear)
        return x2, o_tp(attn2)
    if os.environ.get("ASQ_OPROJ_SPLIT", "1") != "0" and om.act_quant == "per-tensor":
        # per-tensor o_proj: a stand-alone quantisation kernel + the int8-in GEMM instead of the in-kernel prologue.  The
        # attention output comes from a library kernel, so no producer can emit it as int8; inside the GEMM launch the
        # prologue is exposed before the first MMA of a 1.73-round problem (50.7 us), as its own ~5 us launch it is not:
        # 38.0 us for the GEMM, 12.74 vs 13.17 ms per Llama-2-7B step (profiles/r02/ab_oproj_split.md).  Bit-identical
        # (the same row routine quantises in both).  ASQ_OPROJ_SPLIT=0 restores the single launch.
        q8o, _ = _lib.quantize_act(attn2, _lib.ACT_SCALE, float(om.quant_scale.item()))
        y = _lib.w8a8_linear_q8(q8o, om.weight, om.bias if om.use_bias else None, float(om.dequant_scale.item()),
                                out_dtype=x2.dtype, residual=x2 if fuse_res else None)
        return (y, None) if fuse_res else (x2, y)
    if fuse_res:
        # residual add in o_proj's epilogue: x2 <- T(x2 + o_proj(attn)); the norm kernel then reads one tensor
        if om.act_quant == "per-token":
            mode, qs = _lib.ACT_PER_TOKEN, 1.0
        else:
            mode, qs = _lib.ACT_SCALE, float(om.quant_scale.item())
        x2 = _lib.w8a8_linear(attn2, om.weight, om.bias if om.use_bias else None, mode, qs, float(om.dequant_scale.item()),
                              residual=x2)
        return x2, None
    return x2, om(attn2)


def _mlp_half_glue(layer: "QuantDecoderLayer", x2: torch.Tensor, delta: Optional[torch.Tensor]):
    """Second half of a Llama block: add+RMSNorm -> int8, gate|up GEMM with SiLU(gate)*up (and down_proj's
    quantisation when it is per-tensor) in the epilogue, down_proj with the residual add in its epilogue (one GPU)
    or the row-parallel all-reduce behind it (tensor parallel)."""
    from . import _lib

    cfg = layer.cfg
    tp_world = getattr(layer, "tp_world", 1)
    fuse_res = tp_world == 1 and os.environ.get("ASQ_RESIDUAL_EPILOGUE", "1") != "0"
    norm_prologue = fuse_res and os.environ.get("ASQ_NORM_PROLOGUE", "0") == "1"
    down, down_tp = _row_parallel(layer, "down_proj")
    il = getattr(layer, "gate_up_il", None)
    per_token_down = down.act_quant == "per-token"
    out_qs = None if per_token_down else float(down.quant_scale.item())
    if norm_prologue and delta is None and il is not None:
        a = _lib.w8a8_rmsnorm_gateup_swiglu(x2, layer.post_attention_layernorm_weight, cfg.rms_eps, il[0], il[1], il[2],
                                            up_dequant_scale=il[3], out_quant_scale=out_qs)
    else:
        x2, _, q8 = _lib.add_rmsnorm_quant(x2, delta, layer.post_attention_layernorm_weight, cfg.rms_eps)
        if il is not None:
            # SiLU(gate)*up — and down_proj's activation quantisation when it is per-tensor — in the gate|up epilogue.
            # per-token fc2: the row absmax needs the whole SiLU*up row, so the epilogue emits the product in the
            # activation dtype and down_proj quantises it per token in its own launch
            a = _lib.w8a8_gateup_swiglu(q8, il[0], il[1], il[2], up_dequant_scale=il[3], out_quant_scale=out_qs, mid_dtype=x2.dtype)
        else:
            gu_mod = layer.gate_up_proj
            gu = _lib.w8a8_linear_q8(q8, gu_mod.weight, gu_mod.bias if gu_mod.use_bias else None, 1.0,
                                     col_scale=gu_mod._col_scale(x2.device), out_dtype=x2.dtype)
            if per_token_down:
                _, a = _lib.silu_mul_quant(gu, 1.0, want_q=False, want_a=True)
            else:
                a, _ = _lib.silu_mul_quant(gu, out_qs)
    if down_tp is not None:
        return x2, (down_tp(a) if per_token_down else down_tp.forward_q8(a, out_dtype=x2.dtype))
    bias = down.bias if down.use_bias else None
    ds = float(down.dequant_scale.item())
    if per_token_down:
        if fuse_res:
            return _lib.w8a8_linear(a, down.weight, bias, _lib.ACT_PER_TOKEN, 1.0, ds, residual=x2), None
        return x2, down(a)
    # x2 <- T(x2 + down_proj(a)) in the epilogue; nothing is left to add in the next norm kernel
    y = _lib.w8a8_linear_q8(a, down.weight, bias, ds, out_dtype=x2.dtype, residual=x2 if fuse_res else None)
    return (y, None) if fuse_res else (x2, y)


def _moe_half_glue(layer: "QuantDecoderLayer", x2: torch.Tensor, delta: Optional[torch.Tensor]):
    """Second half of a Mixtral block: add+RMSNorm (the activation-dtype output feeds the router and the per-token
    experts), routing in torch, two grouped launches for all experts, one all-reduce under tensor parallelism."""
    from . import _lib

    x2, h, _ = _lib.add_rmsnorm_quant(x2, delta, layer.post_attention_layernorm_weight, layer.cfg.rms_eps, want_h=True, want_q=False)
    return x2, layer._moe_forward(h)


_NVTX = os.environ.get("ASQ_NVTX", "0") == "1"  # named ranges per decoder-block half for timeline tools (off: zero cost)


class _Range:
    """`with _Range("layer3.attn"):` -> an NVTX range when ASQ_NVTX=1 (host-side markers only: legal under graph capture)."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def _layer_forward_glue(layer: "QuantDecoderLayer", x2: torch.Tensor, delta: Optional[torch.Tensor], B: int, S: int,
                        rope: RopeTables, index: int = 0):
    with _Range(f"asq.layer{index}.attention"):
        x2, delta = _attention_half_glue(layer, x2, delta, B, S, rope)
    with _Range(f"asq.layer{index}.mlp"):
        if layer.moe is not None:
            return _moe_half_glue(layer, x2, delta)
        return _mlp_half_glue(layer, x2, delta)


class QuantDecoder(nn.Module):
    """Embedding -> N quantized decoder layers -> norm -> lm_head (bf16, not quantized, as in the reference).
    tp = (rank, world) builds the tensor-parallel shard of every layer as it is created (the full-size layer of
    a 70B / 8x7B model never accumulates); ``tp.build_tp_decoder`` then wraps the row-parallel projections."""

    def __init__(self, cfg: DecoderConfig, quant_config: Optional[Dict[str, str]] = None, device="cuda",
                 dtype=torch.bfloat16, seed: int = 0, layers: Optional[int] = None, fuse_projections: bool = False,
                 glue: bool = False, swiglu_epilogue: bool = True, tp=None):
        super().__init__()
        self.cfg = cfg
        self.qcfg = normalise_quant_config(quant_config or {})
        # producer-side fusions need the fused projections and per-tensor qkv (the norm emits its int8 input); the
        # dense MLP additionally needs per-tensor fc1.  out / fc2 may be per-token (BASELINE configs 3, 4): those
        # linears then quantise inside their own launch
        need = ("qkv",) if cfg.experts else ("qkv", "fc1")
        self.glue = bool(glue and fuse_projections and self.qcfg["type"] == "int8"
                         and all(self.qcfg[k] == "per-tensor" for k in need))
        self.dtype = dtype
        gen = torch.Generator(device=device).manual_seed(seed)
        n_layers = cfg.layers if layers is None else layers
        self.embed = nn.Embedding(cfg.vocab, cfg.hidden, device=device, dtype=dtype)
        with torch.no_grad():
            self.embed.weight.normal_(0.0, 1.0, generator=gen)
        self.layers = nn.ModuleList()
        for _ in range(n_layers):
            layer = QuantDecoderLayer(cfg, self.qcfg, gen, device, dtype, fuse_projections, tp=tp)
            if self.glue and swiglu_epilogue and layer.moe is None and layer.gate_up_proj.weight.shape[0] % 64 == 0:
                layer.enable_swiglu_epilogue()
            self.layers.append(layer)
        self.register_buffer("norm_weight", torch.ones(cfg.hidden, dtype=dtype, device=device))
        self.lm_head = nn.Linear(cfg.hidden, cfg.vocab, bias=False, device=device, dtype=dtype)
        with torch.no_grad():
            self.lm_head.weight.normal_(0.0, 0.02, generator=gen)
        self._rope = None

    def quantized_linears(self):
        return [m for layer in self.layers for m in layer.linears()]

    def rope_tables(self, S: int, device) -> RopeTables:
        if self._rope is None or self._rope.seq != S or self._rope.cos.device != device:
            self._rope = RopeTables(S, self.cfg.head_dim, self.cfg.rope_theta, device, self.dtype)
        return self._rope

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, last_token_only: bool = True) -> torch.Tensor:
        B, S = input_ids.shape
        rope = self.rope_tables(S, input_ids.device)
        x = self.embed(input_ids)
        if self.glue:
            from . import _lib

            x2, delta = x.view(B * S, -1), None
            for i, layer in enumerate(self.layers):
                x2, delta = _layer_forward_glue(layer, x2, delta, B, S, rope, i)
            if last_token_only:  # only the last position of every sequence feeds the lm_head
                x2 = x2.view(B, S, -1)[:, -1, :].contiguous()
                delta = delta.view(B, S, -1)[:, -1, :].contiguous() if delta is not None else None
            _, h, _ = _lib.add_rmsnorm_quant(x2, delta, self.norm_weight, self.cfg.rms_eps, want_h=True, want_q=False)
            return self.lm_head(h.view(B, -1, self.cfg.hidden)).float()
        for layer in self.layers:
            x = layer(x, rope.cos, rope.sin)
        if last_token_only:
            x = x[:, -1:, :]
        x = rms_norm_hf(x, self.norm_weight, self.cfg.rms_eps)
        return self.lm_head(x).float()


def quantized_linear_ops(cfg: DecoderConfig, tokens: int, layers: Optional[int] = None) -> float:
    """Algorithmic INT8 ops (2*MACs) of the quantized linears for `tokens` tokens."""
    n_layers = cfg.layers if layers is None else layers
    return 2.0 * cfg.linear_macs_per_token_per_layer() * n_layers * tokens
