"""Tensor-parallel path on real GPUs (needs >= 2 CUDA devices; skipped on the single-GPU test box).

world_size 2 over NCCL: the row-parallel int32 mode must reproduce the unsharded module bit for bit
(per-tensor and per-token, the latter through the max-all-reduced row scales), and the tensor-parallel
decoder stack must agree with the single-GPU stack up to the bf16 rounding of the partial sums.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from autosmoothquant_b200 import harness, tp
        from autosmoothquant_b200.layers.nn.linear import W8A8BFP32OFP32LinearWithQuantScale

        torch.manual_seed(0)
        K, N = 1024, 768
        lin = torch.nn.Linear(K, N, bias=True)
        x = torch.randn(300, K).to(torch.bfloat16).to(dev)
        results = {}
        for act in ("per-tensor", "per-token"):
            full = W8A8BFP32OFP32LinearWithQuantScale.from_float(lin, 0.05, act_quant=act).to(dev)
            want = full(x)
            row = tp.RowParallelLinear(tp.shard_row(full, rank, world).to(dev), reduce="int32", has_bias=True)
            lo, hi = rank * K // world, (rank + 1) * K // world
            got = row(x[:, lo:hi].contiguous())
            results[f"int32 {act}"] = bool(torch.equal(got, want))
            row_n = tp.RowParallelLinear(tp.shard_row(full, rank, world).to(dev), reduce="native", has_bias=True)
            got_n = row_n(x[:, lo:hi].contiguous())
            results[f"native {act}"] = bool(torch.allclose(got_n.float(), want.float(), rtol=2 ** -6, atol=2 ** -6 * float(want.abs().max())))
        # FP8-e4m3 per-token (BASELINE config 5): column- then row-parallel pair against the unsharded modules
        from autosmoothquant_b200.layers.nn.linear import FP8LinearDynamic
        lin1, lin2 = torch.nn.Linear(K, 1536, bias=False), torch.nn.Linear(1536, N, bias=True)
        f1 = FP8LinearDynamic.from_float(lin1, act_quant="per-token", reference_compat=False)._apply(lambda t: t.to(dev))
        f2 = FP8LinearDynamic.from_float(lin2, act_quant="per-token", reference_compat=False)._apply(lambda t: t.to(dev))
        xf = x.float()  # the reference's fp8 path effectively runs in fp32
        want = f2(f1(xf))
        col = tp.ColumnParallelLinear(tp.shard_column(f1, rank, world)._apply(lambda t: t.to(dev)))
        row = tp.RowParallelLinear(tp.shard_row(f2, rank, world)._apply(lambda t: t.to(dev)), reduce="fp32", has_bias=True)
        h_local = col(xf)
        lo, hi = rank * 1536 // world, (rank + 1) * 1536 // world
        results["fp8 column shard == slice of the unsharded output"] = bool(torch.equal(h_local, f1(xf)[:, lo:hi]))
        got = row(h_local)
        results["fp8 row-parallel per-token"] = bool(torch.allclose(got, want, rtol=1e-4, atol=1e-4 * float(want.abs().max())))
        # decoder stacks: tensor parallel vs single GPU.  Dense INT8 in the exact modes must be BIT-EQUAL to the unsharded
        # stack (this is bench.py's tp_parity gate); the rounded-partial modes, FP8 and the MoE stack agree within the
        # bf16 rounding of the partial sums.
        def close(a, b):
            return bool(float((a - b).abs().max()) <= 0.08 * float(b.abs().max()))

        ids = torch.randint(0, harness.TINY.vocab, (2, 64), generator=torch.Generator().manual_seed(0)).to(dev)
        for cfg_name in ("tiny", "tiny-gqa"):
            cfg = harness.CONFIGS[cfg_name]
            ref = harness.QuantDecoder(cfg, {}, device=dev, seed=3, fuse_projections=True, glue=True)(ids, last_token_only=False)
            for mode in ("nccl", "nccl-int32", "fused", "fused-int32", "nvls"):
                model = tp.build_tp_decoder(cfg, layers=None, device=dev, world=world, rank=rank, seed=3, glue=True,
                                            tp_reduce=mode, max_tokens=128)
                out = model(ids, last_token_only=False)
                if mode.endswith("int32"):
                    results[f"{cfg_name} decoder {mode} bit-equal to the unsharded stack"] = bool(torch.equal(out, ref))
                else:
                    results[f"{cfg_name} decoder {mode}"] = close(out, ref)
                results[f"{cfg_name} decoder {mode} repeatable"] = bool(torch.equal(out, model(ids, last_token_only=False)))
                if mode == "fused":  # 16-bit partials reproduce GEMM + bf16 NCCL all-reduce exactly at world 2
                    ref_nccl = tp.build_tp_decoder(cfg, layers=None, device=dev, world=world, rank=rank, seed=3, glue=True)(ids, last_token_only=False)
                    results[f"{cfg_name} decoder fused == decoder NCCL"] = bool(torch.equal(out, ref_nccl))
                if model.peer_comm is not None:
                    model.peer_comm.close()
            model = tp.build_tp_decoder(cfg, layers=None, device=dev, world=world, rank=rank, seed=3, glue=False)
            results[f"{cfg_name} decoder module path (glue=False)"] = close(model(ids, last_token_only=False), ref)
        # BASELINE config 3 / 4 / 5 in miniature: per-token out / fc2, the Mixtral stack, the FP8 per-token stack
        pt = {"out": "per-token", "fc2": "per-token"}
        fp8 = {"type": "fp8", "qkv": "per-token", "out": "per-token", "fc1": "per-token", "fc2": "per-token"}
        for cfg_name, qc, modes in (("tiny-gqa", pt, ("nccl", "nccl-int32", "fused-int32", "nvls")),
                                    ("tiny-moe", {"fc1": "per-token", "fc2": "per-token"}, ("nccl", "nvls")),
                                    ("tiny-moe", {}, ("nccl",)),
                                    ("tiny-gqa", fp8, ("nccl", "nvls"))):
            cfg = harness.CONFIGS[cfg_name]
            ref = harness.QuantDecoder(cfg, qc, device=dev, seed=3, fuse_projections=True, glue=True)(ids, last_token_only=False)
            tag = f"{cfg_name} {qc.get('type', 'int8')} {'per-token' if qc else 'per-tensor'}"
            for mode in modes:
                model = tp.build_tp_decoder(cfg, layers=None, device=dev, world=world, rank=rank, quant_config=qc, seed=3,
                                            glue=True, tp_reduce=mode, max_tokens=128)
                out = model(ids, last_token_only=False)
                if mode.endswith("int32"):
                    results[f"{tag} decoder {mode} bit-equal to the unsharded stack"] = bool(torch.equal(out, ref))
                else:
                    results[f"{tag} decoder {mode}"] = close(out, ref)
                if model.peer_comm is not None:
                    model.peer_comm.close()
        ret[rank] = results
    finally:
        dist.destroy_process_group()


def test_tp2_nccl_matches_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)  # noqa: a hang is bounded by pytest-timeout
    assert len(ret) == world
    for rank in range(world):
        for name, ok in ret[rank].items():
            assert ok, f"rank {rank}: {name}"
