"""Tensor parallelism for the quantized linears (Megatron layout, one process per GPU).

The reference has no parallelism at all; BASELINE configs 4-5 shard qkv / fc1
column-wise (split N: replicated input, no communication) and out / fc2 row-wise (split K: each
rank multiplies its slice of the activation by its slice of the weight, partial [M,N] products are
summed with ONE all-reduce over NCCL / NVLink).  The weight scale is per-tensor, so every shard
keeps the full module's ``dequant_scale``; the fp32 bias is added by rank 0's shard only.

Row-parallel + per-token: the reference semantics need the row absmax over the WHOLE K, so the
local per-token scales (monotonic in the local absmax) are max-all-reduced first and the kernel
quantises with the supplied scales (``ASQ_ACT_ROW_SCALE_GIVEN``); ``local_scales=True`` skips that
exchange (each rank dequantises with its own scale: valid, slightly more accurate, not bit-identical
to the single-GPU reference).

``reduce="int32"`` is the exactness mode used by the tests: the int32 accumulators are all-reduced
and dequantised afterwards, which is bit-identical to the unsharded module.

Compute is injectable (``backend``) so the host-side logic is testable on CPU with gloo and the
oracle; the default backend is the CUDA library and has no CPU fallback.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist
from torch import nn

from . import _lib
from .layers.nn.linear import (
    FP8LinearDynamic,
    W8A8BFP32OFP32Linear,
    W8A8BFP32OFP32LinearWithQuantScale,
)


def _shard_bounds(size: int, rank: int, world: int, align: int = 16):
    if size % (world * align) != 0:
        raise ValueError(f"dimension {size} is not divisible by world {world} x {align}")
    step = size // world
    return rank * step, (rank + 1) * step


def _shard_fp8(module: FP8LinearDynamic, rows: slice, cols: slice, keep_bias: bool) -> FP8LinearDynamic:
    """FP8-e4m3 dynamic module (BASELINE config 5) restricted to weight[rows, cols]; the per-tensor weight scale is
    shared by every shard, exactly as the INT8 dequant scale."""
    w = module.weight.view(torch.uint8)[rows, cols].contiguous().view(torch.float8_e4m3fn)
    out = FP8LinearDynamic(w.shape[1], w.shape[0], module.act_quant, keep_bias)
    out.weight = w
    out.weight_scale = module.weight_scale.clone()
    if keep_bias:
        out.bias = module.bias[rows].clone()
    return out


def shard_column(module: nn.Module, rank: int, world: int) -> nn.Module:
    """Rank's column-parallel shard (rows [rN/p, (r+1)N/p) of the [N,K] weight) of an INT8 / FP8 module."""
    lo, hi = _shard_bounds(module.out_features, rank, world, align=1)
    if isinstance(module, FP8LinearDynamic):
        return _shard_fp8(module, slice(lo, hi), slice(None), module.use_bias)
    out = type(module)(module.in_features, hi - lo, module.use_bias, module.act_quant)
    out.weight = module.weight[lo:hi].contiguous()
    if module.use_bias:
        out.bias = module.bias[lo:hi].contiguous()
    for name in module._scale_names:
        setattr(out, name, getattr(module, name).clone())
    return out


def shard_row(module: nn.Module, rank: int, world: int) -> nn.Module:
    """Rank's row-parallel shard (columns [rK/p, (r+1)K/p) of the [N,K] weight); bias stays on rank 0."""
    lo, hi = _shard_bounds(module.in_features, rank, world, align=16)
    keep_bias = module.use_bias and rank == 0
    if isinstance(module, FP8LinearDynamic):
        return _shard_fp8(module, slice(None), slice(lo, hi), keep_bias)
    out = type(module)(hi - lo, module.out_features, keep_bias, module.act_quant)
    out.weight = module.weight[:, lo:hi].contiguous()
    if keep_bias:
        out.bias = module.bias.clone()
    for name in module._scale_names:
        setattr(out, name, getattr(module, name).clone())
    return out


# ------------------------------------------------------------------------------ compute backends
class CudaBackend:
    """The product path: fused kernels through the C ABI."""

    @staticmethod
    def linear(module, x2, mode, quant_scale, row_scale=None):
        return _lib.w8a8_linear(x2, module.weight, module.bias if module.use_bias else None, mode, quant_scale,
                                float(module.dequant_scale.item()), row_scale_out=row_scale)

    @staticmethod
    def local_row_scales(x2):
        _, s = _lib.quantize_act(x2, _lib.ACT_PER_TOKEN)
        return s

    @staticmethod
    def fp8_local_row_scales(x2):
        _, s = _lib.quantize_act(x2, _lib.ACT_PER_TOKEN, fp8=True)
        return s

    @staticmethod
    def fp8_linear_given_scales(module, x2, row_scale, out_fp32):
        return _lib.fp8_linear(x2, module.weight, module._bias_f32(), _lib.ACT_ROW_SCALE_GIVEN, 1.0,
                               float(module.weight_scale.item()), out_dtype=torch.float32 if out_fp32 else None,
                               row_scale_out=row_scale)

    @staticmethod
    def int32_partial(module, x2, mode, quant_scale, row_scale=None):
        if mode == _lib.ACT_PER_TOKEN:
            raise NotImplementedError("int32 reduction needs global row scales (local_scales=False)")
        q, _ = _lib.quantize_act(x2, mode, quant_scale, row_scale=row_scale)
        acc = torch.empty((x2.shape[0], module.out_features), dtype=torch.int32, device=x2.device)
        _lib.i8gemm_o32(q, module.weight, acc)
        return acc


class ColumnParallelLinear(nn.Module):
    """qkv / fc1: every rank holds N/p output features; input replicated, output stays sharded."""

    def __init__(self, shard: nn.Module, backend=CudaBackend):
        super().__init__()
        self.shard = shard
        self.backend = backend
        self.in_features, self.out_features = shard.in_features, shard.out_features

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        m = self.shard
        if isinstance(m, FP8LinearDynamic):  # replicated input: the module's own fused launch, output stays sharded
            if hasattr(self.backend, "fp8_module_forward"):
                return self.backend.fp8_module_forward(m, x)
            return m(x)
        x2 = x.reshape(-1, m.in_features)
        if m.act_quant == "per-token":
            mode, qs = _lib.ACT_PER_TOKEN, 1.0
        elif isinstance(m, W8A8BFP32OFP32LinearWithQuantScale):
            mode, qs = _lib.ACT_SCALE, float(m.quant_scale.item())
        else:
            mode, qs = _lib.ACT_ROUND, 1.0
        y = self.backend.linear(m, x2, mode, qs)
        return y.view(*x.shape[:-1], m.out_features)


class RowParallelLinear(nn.Module):
    """out / fc2: every rank holds K/p input features; partial outputs are all-reduced (sum).

    ``has_bias`` tells every rank whether the unsharded module had a bias (only rank 0's shard keeps it).
    """

    def __init__(self, shard: nn.Module, group=None, reduce: str = "native", local_scales: bool = False,
                 backend=CudaBackend, has_bias: Optional[bool] = None, comm=None):
        super().__init__()
        if reduce not in ("native", "fp32", "int32", "fused", "fused-native"):
            raise ValueError("reduce must be 'native' (activation dtype), 'fp32', 'int32', 'fused' or 'fused-native'")
        if reduce.startswith("fused") and comm is None:
            raise ValueError("reduce='fused' needs a peer.PeerComm (GEMM + all-reduce in one kernel over peer memory)")
        self.comm = comm
        self.shard, self.group, self.reduce, self.local_scales, self.backend = shard, group, reduce, local_scales, backend
        self.has_bias = shard.use_bias if has_bias is None else has_bias
        self.in_features, self.out_features = shard.in_features, shard.out_features
        self._full_bias = None

    def _bias_everywhere(self, device) -> Optional[torch.Tensor]:
        """int32 mode dequantises after the reduction on every rank, so every rank needs the bias."""
        if not self.has_bias:
            return None
        if self._full_bias is None:
            b = (self.shard.bias.clone().to(device) if self.shard.use_bias
                 else torch.zeros(self.out_features, dtype=torch.float32, device=device))
            dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)  # only rank 0 contributes non-zeros
            self._full_bias = b
        return self._full_bias

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        m = self.shard
        x2 = x.reshape(-1, m.in_features)
        if isinstance(m, FP8LinearDynamic):
            return self._forward_fp8(m, x, x2)
        row_scale = None
        if m.act_quant == "per-token":
            if self.local_scales:
                mode, qs = _lib.ACT_PER_TOKEN, 1.0
            else:
                row_scale = self.backend.local_row_scales(x2)
                dist.all_reduce(row_scale, op=dist.ReduceOp.MAX, group=self.group)  # = scale of the global row absmax
                mode, qs = _lib.ACT_ROW_SCALE_GIVEN, 1.0
        elif isinstance(m, W8A8BFP32OFP32LinearWithQuantScale):
            mode, qs = _lib.ACT_SCALE, float(m.quant_scale.item())
        else:
            mode, qs = _lib.ACT_ROUND, 1.0

        if self.reduce.startswith("fused"):
            # ONE launch per rank, no NCCL (peer.PeerComm).  "fused": int32 partials over NVLink peer stores, exact
            # owner-side sum, the unsharded module's fp32 epilogue -> bit-identical to the unsharded module.
            # "fused-native": 16-bit dequantised partials (half the bytes) = the numerics of reduce="native".
            if mode == _lib.ACT_PER_TOKEN:
                raise NotImplementedError("reduce='fused' needs global row scales (local_scales=False)")
            q, _ = _lib.quantize_act(x2, mode, qs, row_scale=row_scale)
            if self.reduce == "fused":
                y = self.comm.linear_q8_allreduce(q, m.weight, self._bias_everywhere(x2.device),
                                                  float(m.dequant_scale.item()), row_scale=row_scale)
            else:
                y = self.comm.linear_q8_allreduce(q, m.weight, m.bias if m.use_bias else None,
                                                  float(m.dequant_scale.item()), row_scale=row_scale, partials="native")
            return y.view(*x.shape[:-1], m.out_features)

        if self.reduce == "int32":
            # exactness mode: sum the integer accumulators, then the unsharded module's fp32 epilogue
            acc = self.backend.int32_partial(m, x2, mode, qs, row_scale)
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
            ds = float(m.dequant_scale.item())
            y = (ds * row_scale.view(-1, 1)) * acc if row_scale is not None else ds * acc
            bias = self._bias_everywhere(y.device)
            if bias is not None:
                y = y + bias
            return y.to(x.dtype).view(*x.shape[:-1], m.out_features)

        y = self.backend.linear(m, x2, mode, qs, row_scale)
        if self.reduce == "fp32" and y.dtype != torch.float32:
            y = y.float()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y.to(x.dtype).view(*x.shape[:-1], m.out_features)


def _rowparallel_forward_fp8(self, m: FP8LinearDynamic, x: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """FP8 per-token row-parallel (config 5): the per-token scale is absmax(row over the WHOLE K) / 448, so the local
    scales are max-all-reduced and the kernel quantises with the supplied scales; fp32-accumulated partial
    products are rounded to the activation dtype (or kept fp32, reduce="fp32") and summed by one all-reduce."""
    if m.act_quant != "per-token" or self.reduce in ("int32", "fused"):
        raise NotImplementedError("FP8 row-parallel: per-token activations with reduce='native' or 'fp32' only")
    row_scale = self.backend.fp8_local_row_scales(x2)
    if not self.local_scales:
        dist.all_reduce(row_scale, op=dist.ReduceOp.MAX, group=self.group)
    y = self.backend.fp8_linear_given_scales(m, x2, row_scale, self.reduce == "fp32")
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
    return y.to(x.dtype).view(*x.shape[:-1], m.out_features)


RowParallelLinear._forward_fp8 = _rowparallel_forward_fp8


def _shard_fused_columns(fused, rank: int, world: int, device):
    """Column shard of a fused W_pack module (q|k|v or gate|up): every block is split separately so rank r
    holds [q_r; k_r; v_r] (its own heads) and the per-block dequant scales stay valid."""
    from .layers.nn.linear import W8A8BFP32OFP32QKVLinear

    sizes = fused.qkv_size
    local_sizes, rows = [], []
    start = 0
    for n in sizes:
        if n % world:
            raise ValueError(f"block of {n} output features is not divisible by world {world}")
        step = n // world
        rows.append(fused.weight[start + rank * step:start + (rank + 1) * step])
        local_sizes.append(step)
        start += n
    out = W8A8BFP32OFP32QKVLinear(local_sizes, fused.in_features, sum(local_sizes), fused.use_bias, fused.act_quant)
    out.weight = torch.cat(rows, dim=0).contiguous()
    for name in fused._scale_names:
        setattr(out, name, getattr(fused, name).clone())
    if fused.use_bias:
        start, parts = 0, []
        for n in sizes:
            step = n // world
            parts.append(fused.bias[start + rank * step:start + (rank + 1) * step])
            start += n
        out.bias = torch.cat(parts).contiguous()
    return out.to(device)


def build_tp_decoder(cfg, layers, device, world: int, rank: int, quant_config: Optional[Dict[str, str]] = None,
                     group=None, dtype=torch.bfloat16, seed: int = 0, glue: bool = True, fused_allreduce: bool = False,
                     max_tokens: int = 2048, partials: str = "int32"):
    """The benchmark stack of ``harness.QuantDecoder`` tensor-parallel over `world` ranks: fused q|k|v and
    gate|up column-sharded by head / by intermediate channel, o_proj and down_proj row-sharded with ONE
    all-reduce each (NCCL over NVLink), attention over the local heads, residual stream and norms replicated.
    Every rank builds the same seeded full-size layer and keeps its shard (one layer at a time).
    fused_allreduce=True replaces (GEMM launch + NCCL all-reduce) of the two row-parallel projections by the
    single-launch peer-memory kernel of ``peer.PeerComm`` (``max_tokens`` = largest batch*seq it will see)."""
    from . import harness

    model = harness.QuantDecoder(cfg, quant_config, device=device, dtype=dtype, seed=seed, layers=layers,
                                 fuse_projections=True, glue=glue, swiglu_epilogue=False)
    if model.qcfg["type"] != "int8":
        raise NotImplementedError("tensor-parallel fp8 stack")
    comm = None
    if fused_allreduce and world > 1:
        from .peer import PeerComm

        comm = PeerComm(group=group, device=device, max_m=max_tokens, max_n=cfg.hidden, dtype=dtype)
    model.peer_comm = comm
    for layer in model.layers:
        layer.qkv_proj = _shard_fused_columns(layer.qkv_proj, rank, world, device)
        layer.qkv_sizes = list(layer.qkv_proj.qkv_size)
        layer.gate_up_proj = _shard_fused_columns(layer.gate_up_proj, rank, world, device)
        if model.glue and (cfg.intermediate // world) % 32 == 0:
            layer.enable_swiglu_epilogue()  # interleave the LOCAL gate / up rows for the SwiGLU epilogue
        for name in ("o_proj", "down_proj"):
            full = getattr(layer, name)
            setattr(layer, name, RowParallelLinear(shard_row(full, rank, world).to(device), group=group,
                                                   has_bias=full.use_bias,
                                                   reduce=("fused" if partials == "int32" else "fused-native") if comm is not None else "native",
                                                   comm=comm))
        layer.tp_world, layer.tp_group, layer.peer_comm, layer.peer_partials = world, group, comm, partials
        torch.cuda.empty_cache()
    model.tp_world = world
    return model
