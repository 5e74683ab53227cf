"""Tensor-parallel host logic on CPU: world_size 2 over gloo, compute injected from the oracle.

The sharding / all-reduce plumbing of autosmoothquant_b200.tp is exercised without a GPU by swapping
the CUDA backend for one that evaluates the oracle on CPU tensors.  The exactness mode (int32
all-reduce) must reproduce the unsharded oracle bit-for-bit, including per-token activations whose
global row scale needs a max-all-reduce first.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import w8a8_oracle as O

MODE_NAMES = {0: "round", 1: "scale", 2: "per-token", 4: "given"}


class OracleBackend:
    """Same interface as tp.CudaBackend, evaluated with the numpy oracle (test-only)."""

    dtype = "bf16"

    @classmethod
    def _quant(cls, x2, mode, qs, row_scale):
        x = x2.float().numpy()
        if mode == 4:
            with np.errstate(divide="ignore", invalid="ignore"):
                return O.sat_i8(np.rint(x / row_scale.numpy().reshape(-1, 1)))
        return O.quantize_act_int8(x, cls.dtype, MODE_NAMES[mode], qs)[0]

    @classmethod
    def linear(cls, module, x2, mode, quant_scale, row_scale=None):
        q = cls._quant(x2, mode, quant_scale, row_scale)
        if mode == 2:
            row_scale = torch.from_numpy(O.quantize_act_int8(x2.float().numpy(), cls.dtype, "per-token")[1])
        acc = O.int8_gemm_i32(q, module.weight.numpy())
        ds = np.float32(module.dequant_scale.item())
        f = ds * row_scale.numpy().reshape(-1, 1) if row_scale is not None else ds
        y = O._dequant(acc, f, module.bias.numpy() if module.use_bias else None, cls.dtype)
        return torch.from_numpy(y).to(x2.dtype)

    @classmethod
    def local_row_scales(cls, x2):
        return torch.from_numpy(O.quantize_act_int8(x2.float().numpy(), cls.dtype, "per-token")[1])

    @classmethod
    def int32_partial(cls, module, x2, mode, quant_scale, row_scale=None):
        return torch.from_numpy(O.int8_gemm_i32(cls._quant(x2, mode, quant_scale, row_scale), module.weight.numpy()))

    # ---- FP8-e4m3 per-token (BASELINE config 5), fp32 activations as the reference effectively runs it
    @staticmethod
    def _w_bytes(module):
        return module.weight.view(torch.uint8).numpy()

    @classmethod
    def fp8_local_row_scales(cls, x2):
        return torch.from_numpy(O.quantize_act_fp8(x2.float().numpy(), "f32", "per-token")[1])

    @classmethod
    def fp8_linear_given_scales(cls, module, x2, row_scale, out_fp32):
        x = x2.float().numpy()
        s = row_scale.numpy().reshape(-1, 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            q = O.e4m3_encode(np.clip(x / s, -448.0, 448.0))
        y = O.fp8_linear_exact(q, cls._w_bytes(module), s.reshape(-1), float(module.weight_scale),
                               module.bias.numpy() if module.use_bias else None)
        return torch.from_numpy(y.astype(np.float32))

    @classmethod
    def fp8_module_forward(cls, module, x):
        x2 = x.reshape(-1, module.in_features).float().numpy()
        q, s = O.quantize_act_fp8(x2, "f32", "per-token")
        y = O.fp8_linear_exact(q, cls._w_bytes(module), s, float(module.weight_scale), module.bias.numpy() if module.use_bias else None)
        return torch.from_numpy(y.astype(np.float32)).view(*x.shape[:-1], module.out_features)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, act_quant, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from autosmoothquant_b200 import tp
        from autosmoothquant_b200.layers.nn.linear import W8A8BFP32OFP32Linear, W8A8BFP32OFP32LinearWithQuantScale

        torch.manual_seed(0)  # identical "checkpoint" on every rank
        K, I = 64, 96
        fc1 = torch.nn.Linear(K, I, bias=True)
        fc2 = torch.nn.Linear(I, K, bias=True)
        up = W8A8BFP32OFP32Linear.from_float(fc1, 0.03, act_quant=act_quant)
        down = W8A8BFP32OFP32LinearWithQuantScale.from_float(fc2, 0.05, act_quant=act_quant)
        x = (torch.randn(3, 7, K) * (30.0 if act_quant == "per-tensor" else 1.0)).to(torch.bfloat16)

        col = tp.ColumnParallelLinear(tp.shard_column(up, rank, world), backend=OracleBackend)
        row = tp.RowParallelLinear(tp.shard_row(down, rank, world), reduce="int32", backend=OracleBackend,
                                   has_bias=True)
        h_local = col(x)                      # [3,7,I/world]: column-parallel output stays sharded
        y = row(h_local)                      # row-parallel consumes the local slice, all-reduces

        # unsharded oracle
        h_full = O.w8a8_linear(x.float().numpy(), "bf16", up.weight.numpy(), float(up.dequant_scale), act_quant=act_quant,
                               bias=up.bias.numpy())
        lo, hi = rank * I // world, (rank + 1) * I // world
        ok_col = np.array_equal(h_local.float().numpy(), h_full[..., lo:hi])
        y_full = O.w8a8_linear(h_full, "bf16", down.weight.numpy(), float(down.dequant_scale), act_quant=act_quant,
                               bias=down.bias.numpy(),
                               quant_scale=float(down.quant_scale) if act_quant == "per-tensor" else None)
        ok_row = np.array_equal(y.float().numpy(), y_full)
        # native-dtype reduction: close, not bit-identical (partial sums are rounded per rank)
        row_native = tp.RowParallelLinear(tp.shard_row(down, rank, world), reduce="fp32", backend=OracleBackend,
                                          has_bias=True)
        y_n = row_native(h_local).float().numpy()
        ok_native = np.allclose(y_n, y_full, rtol=2 ** -6, atol=2 ** -6 * np.abs(y_full).max())
        ret[rank] = (ok_col, ok_row, ok_native)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("act_quant", ["per-tensor", "per-token"])
def test_tp2_matches_unsharded_oracle(act_quant):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, act_quant, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        ok_col, ok_row, ok_native = ret[rank]
        assert ok_col, f"rank {rank}: column-parallel shard differs"
        assert ok_row, f"rank {rank}: int32 row-parallel result differs from the unsharded oracle"
        assert ok_native, f"rank {rank}: fp32-reduced row-parallel result out of tolerance"


def _worker_fp8(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from autosmoothquant_b200 import tp
        from autosmoothquant_b200.layers.nn.linear import FP8LinearDynamic

        torch.manual_seed(0)
        K, I = 64, 96
        f1 = FP8LinearDynamic.from_float(torch.nn.Linear(K, I, bias=False), act_quant="per-token", reference_compat=False)
        f2 = FP8LinearDynamic.from_float(torch.nn.Linear(I, K, bias=True), act_quant="per-token", reference_compat=False)
        x = torch.randn(5, 4, K)
        col = tp.ColumnParallelLinear(tp.shard_column(f1, rank, world), backend=OracleBackend)
        row = tp.RowParallelLinear(tp.shard_row(f2, rank, world), reduce="fp32", backend=OracleBackend, has_bias=True)
        h_local = col(x)
        y = row(h_local)
        h_full = OracleBackend.fp8_module_forward(f1, x)
        lo, hi = rank * I // world, (rank + 1) * I // world
        ok_col = bool(torch.equal(h_local, h_full[..., lo:hi]))
        y_full = OracleBackend.fp8_module_forward(f2, h_full)
        # the global per-token scale (max-all-reduced) makes every rank quantise exactly as the unsharded module does;
        # only the fp32 summation of the two K halves differs from the fp64 evaluation of the whole K
        ok_row = bool(torch.allclose(y, y_full, rtol=1e-5, atol=1e-5 * float(y_full.abs().max())))
        ret[rank] = (ok_col, ok_row, tuple(tp.shard_row(f2, rank, world).weight.shape), tp.shard_row(f2, rank, world).use_bias)
    finally:
        dist.destroy_process_group()


def test_tp2_fp8_per_token_matches_unsharded_oracle():
    """BASELINE config 5 plumbing (FP8-e4m3 per-token, column- then row-parallel) on CPU over gloo."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_fp8, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        ok_col, ok_row, shape, has_bias = ret[rank]
        assert ok_col and ok_row, f"rank {rank}: col {ok_col} row {ok_row}"
        assert shape == (64, 48) and has_bias == (rank == 0)  # K sharded, bias only on rank 0


def test_row_parallel_int32_identity():
    """Pure-oracle statement of the sharding identity used above."""
    rng = np.random.default_rng(0)
    q = rng.integers(-128, 128, size=(33, 128), dtype=np.int8)
    w = rng.integers(-128, 128, size=(40, 128), dtype=np.int8)
    for world in (2, 4, 8):
        np.testing.assert_array_equal(O.tp_row_parallel_int32(q, w, world), O.int8_gemm_i32(q, w))


def test_shard_bounds_validation():
    from autosmoothquant_b200 import tp
    from autosmoothquant_b200.layers.nn.linear import W8A8BFP32OFP32Linear

    m = W8A8BFP32OFP32Linear(48, 30)
    with pytest.raises(ValueError):
        tp.shard_row(m, 0, 2)  # K/p = 24 is not a multiple of 16 (TMA row pitch)
    with pytest.raises(ValueError):
        tp.shard_column(m, 0, 4)  # 30 % 4 != 0
    s = tp.shard_column(m, 1, 2)
    assert tuple(s.weight.shape) == (15, 48)
