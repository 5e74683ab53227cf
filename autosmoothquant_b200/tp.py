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
        out = _shard_fp8(module, slice(None), slice(lo, hi), keep_bias)
        out.full_has_bias = module.use_bias
        return out
    out = type(module)(hi - lo, module.out_features, keep_bias, module.act_quant)
    out.full_has_bias = module.use_bias  # every rank must know whether the unsharded module adds a bias
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


REDUCE_MODES = ("native", "fp32", "int32", "fused", "fused-native", "nvls")


class RowParallelLinear(nn.Module):
    """out / fc2: every rank holds K/p input features; partial outputs are all-reduced (sum).

    reduce =
      "native"        GEMM launch (partials rounded to the activation dtype) + ncclAllReduce
      "fp32"          the same with fp32 partials
      "int32"         exactness mode: int32 accumulators all-reduced over NCCL, then the unsharded module's fp32
                      epilogue -> bit-identical to the unsharded module
      "fused"         ONE launch, no NCCL (peer.PeerComm): int32 partials by NVLink peer stores, exact owner-side sum
                      -> bit-identical to the unsharded module
      "fused-native"  the same kernel with 16-bit dequantised partials (half the bytes; the numerics of "native")
      "nvls"          ONE launch: 16-bit partials stay in this rank's symmetric buffer, the tile's owner reduces them
                      IN THE NVSWITCH (multimem.ld_reduce) and broadcasts the sums (multimem.st); INT8 and FP8

    ``has_bias`` tells every rank whether the unsharded module had a bias (only rank 0's shard keeps it).
    The fused / nvls modes return a tensor that lives in the communicator's double-buffered output; by default it
    is CLONED so callers may keep it.  ``alias_output=True`` (the benchmark stack, which consumes each result at
    once) returns the zero-copy view, valid until the launch after the next one on the same communicator.
    """

    def __init__(self, shard: nn.Module, group=None, reduce: str = "native", local_scales: bool = False,
                 backend=CudaBackend, has_bias: Optional[bool] = None, comm=None, alias_output: bool = False):
        super().__init__()
        if reduce not in REDUCE_MODES:
            raise ValueError(f"reduce must be one of {REDUCE_MODES}")
        if reduce in ("fused", "fused-native", "nvls") and comm is None:
            raise ValueError(f"reduce={reduce!r} needs a peer.PeerComm (GEMM + all-reduce in one kernel over peer memory)")
        self.comm = comm
        self.shard, self.group, self.reduce, self.local_scales, self.backend = shard, group, reduce, local_scales, backend
        self.alias_output = alias_output
        self.has_bias = getattr(shard, "full_has_bias", shard.use_bias) if has_bias is None else has_bias
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

    def _own(self, y: torch.Tensor) -> torch.Tensor:
        return y if self.alias_output else y.clone()

    def _act_mode(self):
        m = self.shard
        if m.act_quant == "per-token":
            return (_lib.ACT_PER_TOKEN, 1.0) if self.local_scales else (_lib.ACT_ROW_SCALE_GIVEN, 1.0)
        if isinstance(m, W8A8BFP32OFP32LinearWithQuantScale):
            return _lib.ACT_SCALE, float(m.quant_scale.item())
        return _lib.ACT_ROUND, 1.0

    @torch.no_grad()
    def forward_q8(self, q: torch.Tensor, row_scale: Optional[torch.Tensor] = None,
                   out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
        """Row-parallel product of activations that are ALREADY int8 (a fused producer quantised them, or forward()
        did): [M, K/p] int8 -> [M, N] in out_dtype, summed over the ranks.  CUDA only."""
        m = self.shard
        ds = float(m.dequant_scale.item())
        if self.reduce == "fused":
            return self._own(self.comm.linear_q8_allreduce(q, m.weight, self._bias_everywhere(q.device), ds, row_scale=row_scale))
        if self.reduce == "fused-native":
            return self._own(self.comm.linear_q8_allreduce(q, m.weight, m.bias if m.use_bias else None, ds, row_scale=row_scale,
                                                           partials="native"))
        if self.reduce == "nvls":
            return self._own(self.comm.linear_q8_allreduce_nvls(q, m.weight, m.bias if m.use_bias else None, ds, row_scale=row_scale))
        if self.reduce == "int32":
            acc = torch.empty((q.shape[0], m.out_features), dtype=torch.int32, device=q.device)
            _lib.i8gemm_o32(q, m.weight, acc)
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
            return self._dequant_i32(acc, ds, row_scale, out_dtype)
        y = _lib.w8a8_linear_q8(q, m.weight, m.bias if m.use_bias else None, ds, row_scale=row_scale,
                                out_dtype=torch.float32 if self.reduce == "fp32" else out_dtype)
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y.to(out_dtype)

    def _dequant_i32(self, acc, ds, row_scale, out_dtype):
        """The unsharded module's epilogue on summed accumulators (linear.py:93,104: factor, * acc, + bias, cast)."""
        y = (ds * row_scale.view(-1, 1)) * acc if row_scale is not None else ds * acc
        bias = self._bias_everywhere(y.device)
        if bias is not None:
            y = y + bias
        return y.to(out_dtype)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        m = self.shard
        x2 = x.reshape(-1, m.in_features)
        if isinstance(m, FP8LinearDynamic):
            return self._forward_fp8(m, x, x2)
        row_scale = None
        mode, qs = self._act_mode()
        if mode == _lib.ACT_ROW_SCALE_GIVEN:
            row_scale = self.backend.local_row_scales(x2)
            dist.all_reduce(row_scale, op=dist.ReduceOp.MAX, group=self.group)  # = scale of the global row absmax
        out_shape = (*x.shape[:-1], m.out_features)

        if self.reduce in ("fused", "fused-native", "nvls"):
            if mode == _lib.ACT_PER_TOKEN:
                raise NotImplementedError(f"reduce={self.reduce!r} needs global row scales (local_scales=False)")
            q, _ = _lib.quantize_act(x2, mode, qs, row_scale=row_scale)
            return self.forward_q8(q, row_scale, x.dtype).view(out_shape)

        if self.reduce == "int32":
            # exactness mode: sum the integer accumulators, then the unsharded module's fp32 epilogue
            acc = self.backend.int32_partial(m, x2, mode, qs, row_scale)
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
            return self._dequant_i32(acc, float(m.dequant_scale.item()), row_scale, x.dtype).view(out_shape)

        y = self.backend.linear(m, x2, mode, qs, row_scale)
        if self.reduce == "fp32" and y.dtype != torch.float32:
            y = y.float()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y.to(x.dtype).view(out_shape)

    def _forward_fp8(self, m: FP8LinearDynamic, x: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        """FP8 per-token row-parallel (config 5): the per-token scale is absmax(row over the WHOLE K) / 448, so the
        local scales are max-all-reduced and the kernel quantises with the supplied scales; fp32-accumulated partial
        products are rounded to the activation dtype (or kept fp32, reduce="fp32") and summed by one all-reduce, or
        — reduce="nvls" — reduced in the switch by the fused kernel."""
        if m.act_quant != "per-token" or self.reduce in ("int32", "fused", "fused-native"):
            raise NotImplementedError("FP8 row-parallel: per-token activations with reduce='native', 'fp32' or 'nvls'")
        row_scale = self.backend.fp8_local_row_scales(x2)
        if not self.local_scales:
            dist.all_reduce(row_scale, op=dist.ReduceOp.MAX, group=self.group)
        out_shape = (*x.shape[:-1], m.out_features)
        if self.reduce == "nvls":
            q, _ = _lib.quantize_act(x2, _lib.ACT_ROW_SCALE_GIVEN, fp8=True, row_scale=row_scale)
            y = self.comm.linear_q8_allreduce_nvls(q, m.weight, m._bias_f32(), float(m.weight_scale.item()), row_scale=row_scale,
                                                   out_dtype=x.dtype if x.dtype != torch.float32 else torch.bfloat16)
            return self._own(y).to(x.dtype).view(out_shape)
        y = self.backend.fp8_linear_given_scales(m, x2, row_scale, self.reduce == "fp32")
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y.to(x.dtype).view(out_shape)


TP_REDUCE_CHOICES = ("nccl", "nccl-int32", "fused", "fused-int32", "nvls")


def build_tp_decoder(cfg, layers, device, world: int, rank: int, quant_config: Optional[Dict[str, str]] = None,
                     group=None, dtype=torch.bfloat16, seed: int = 0, glue: bool = True, tp_reduce: str = "nccl",
                     max_tokens: int = 2048, fused_allreduce: Optional[bool] = None, partials: Optional[str] = None):
    """The benchmark stack of ``harness.QuantDecoder`` tensor-parallel over `world` ranks: q|k|v and gate|up (or
    every expert's w1 / w3) column-sharded by head / by intermediate channel, o_proj and down_proj (w2) row-sharded
    with ONE all-reduce each, attention over the local heads, residual stream, norms and routing replicated.
    Every rank draws the same seeded full-size weights and keeps its shard, one projection at a time.
    tp_reduce picks the row-parallel reduction:
      "nccl"         GEMM launch + ncclAllReduce in the activation dtype
      "nccl-int32"   int32 accumulators over NCCL: bit-identical to the unsharded stack (parity gate)
      "fused-int32"  one launch, int32 partials over NVLink peer stores: bit-identical to the unsharded stack
      "fused"        the same kernel with 16-bit partials (NCCL-native numerics)
      "nvls"         one launch, in-switch reduction (multimem.ld_reduce / multimem.st); INT8 and FP8
    ``max_tokens`` = largest batch*seq the communicator's buffers will see."""
    from . import harness

    if fused_allreduce is not None:  # round-1 keyword pair
        tp_reduce = ("fused-int32" if (partials or "int32") == "int32" else "fused") if fused_allreduce else "nccl"
    if tp_reduce not in TP_REDUCE_CHOICES:
        raise ValueError(f"tp_reduce must be one of {TP_REDUCE_CHOICES}")
    model = harness.QuantDecoder(cfg, quant_config, device=device, dtype=dtype, seed=seed, layers=layers,
                                 fuse_projections=True, glue=glue, tp=(rank, world))
    fp8 = model.qcfg["type"] != "int8"
    if fp8 and tp_reduce in ("nccl-int32", "fused", "fused-int32"):
        raise NotImplementedError("the FP8 stack reduces over NCCL ('nccl') or in the switch ('nvls'): there is no integer partial")
    comm = None
    if tp_reduce in ("fused", "fused-int32", "nvls") and world > 1:
        from .peer import PeerComm

        comm = PeerComm(group=group, device=device, max_m=max_tokens, max_n=cfg.hidden, dtype=dtype, nvls=tp_reduce == "nvls",
                        p2p=tp_reduce != "nvls")
    model.peer_comm = comm
    reduce = {"nccl": "native", "nccl-int32": "int32", "fused": "fused-native", "fused-int32": "fused", "nvls": "nvls"}[tp_reduce]
    for layer in model.layers:
        rows = ("o_proj",) if layer.moe is not None else ("o_proj", "down_proj")
        for name in rows:
            shard = getattr(layer, name)
            setattr(layer, name, RowParallelLinear(shard, group=group,
                                                   reduce=reduce, comm=comm, alias_output=True))
        if not layer.fused:  # module path (FP8): column-parallel wrappers keep the module forward, outputs stay sharded
            cols = ("q_proj", "k_proj", "v_proj") + (() if layer.moe is not None else ("gate_proj", "up_proj"))
            for name in cols:
                setattr(layer, name, ColumnParallelLinear(getattr(layer, name)))
        layer.tp_world, layer.tp_group, layer.peer_comm = world, group, comm
    model.tp_world = world
    model.tp_reduce = tp_reduce
    return model
