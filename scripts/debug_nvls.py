"""NVLS diagnostics (torchrun --nproc-per-node W scripts/debug_nvls.py): (1) what the switch data path sustains for
16-byte multimem.ld_reduce / multimem.st as a function of CTAs x threads x requests in flight (asq_nvls_probe);
(2) element-level comparison of the fused NVLS kernel with GEMM + ncclAllReduce and with the locally evaluated
fp32 sum of the gathered bf16 partials."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L  # noqa: E402
from autosmoothquant_b200 import peer  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
lib = L.load()


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


if "probe" in sys.argv or len(sys.argv) == 1:
    total = int(os.environ.get("PROBE_MB", "128")) << 20          # message size (all ranks' slices together)
    buf = symm.empty(2 * total, dtype=torch.uint8, device=dev)
    buf.view(torch.bfloat16).fill_(1.0)
    hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
    mc = int(hdl.multicast_ptr)
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    slice_b = total // world
    stream = torch.cuda.current_stream().cuda_stream
    dist.barrier()
    rows = []
    for mode, name in ((0, "ld_reduce+st"), (1, "ld_reduce"), (2, "st")):
        for ctas, threads, unroll in ((16, 512, 8), (32, 512, 8), (74, 512, 8), (148, 256, 1), (148, 256, 4), (148, 256, 8), (148, 256, 16),
                                      (148, 512, 8), (148, 1024, 4), (296, 512, 8), (592, 256, 8)):
            def run():
                rc = lib.asq_nvls_probe(ctypes.c_void_p(mc + rank * slice_b), ctypes.c_void_p(mc + total + rank * slice_b), slice_b,
                                        ctas, threads, unroll, mode, ctypes.c_void_p(sink.data_ptr()), ctypes.c_void_p(stream))
                assert rc == 0, lib.asq_last_error()
            us = timeit(run)
            rows.append((name, ctas, threads, unroll, us, slice_b / us / 1e3))
            if rank == 0:
                print(f"probe world {world} {total >> 20} MB message, {name:13s} ctas {ctas:4d} x {threads:4d} thr, {unroll:2d} in flight: "
                      f"{us:8.1f} us  {slice_b / us / 1e3:7.1f} GB/s reduced per GPU  (allreduce algbw {total / us / 1e3:7.1f} GB/s)", flush=True)
    y = buf[total:].view(torch.bfloat16)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("probe result check (ones summed over ranks):", float(y.float().min()), float(y.float().max()), "expected", float(world))
    t = torch.empty(total // 2, dtype=torch.bfloat16, device=dev)
    us = timeit(lambda: dist.all_reduce(t))
    if rank == 0:
        print(f"ncclAllReduce bf16 {total >> 20} MB: {us:8.1f} us  algbw {total / us / 1e3:7.1f} GB/s", flush=True)
    del buf, hdl

if "check" in sys.argv or len(sys.argv) == 1:
    comm = peer.PeerComm(device=dev, max_m=4096, max_n=4096, nvls=True, p2p=False)
    g = torch.Generator().manual_seed(5)
    for (M, N, K) in [(300, 768, 1024), (2048, 4096, 4096), (64, 512, 2048), (256, 256, 512)]:
        a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g).to(dev)
        w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=g).to(dev)
        step = K // world // 16 * 16
        lo, hi = rank * step, (K if rank == world - 1 else (rank + 1) * step)
        al, wl = a[:, lo:hi].contiguous(), w[:, lo:hi].contiguous()
        part = L.w8a8_linear_q8(al, wl, None, 3e-5)
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        local_sum = sum(p.float() for p in parts).to(torch.bfloat16)   # fp32 sum in rank order, one rounding
        y_nccl = part.clone()
        dist.all_reduce(y_nccl)
        for rep in range(3):
            got = comm.linear_q8_allreduce_nvls(al, wl, None, 3e-5).clone()
            torch.cuda.synchronize()
            bad = (got != local_sum)
            nb = int(bad.sum())
            msg = f"rank {rank} {M}x{N}x{K} #{rep}: != fp32-sum {nb}/{got.numel()}, != nccl {int((got != y_nccl).sum())}, nccl != fp32-sum {int((y_nccl != local_sum).sum())}"
            if nb:
                idx = bad.nonzero()
                d = (got.float() - local_sum.float()).abs()
                rel = float((d / (local_sum.float().abs() + 1e-6))[bad].max())
                rows_bad = idx[:, 0].unique()
                cols_bad = idx[:, 1].unique()
                own = (got == parts[rank]) & bad
                other = (got == parts[1 - rank]) & bad if world == 2 else torch.zeros_like(bad)
                msg += (f" | max abs {float(d.max()):.4g} max rel {rel:.3g} | rows {rows_bad[:6].tolist()}..{int(rows_bad[-1])} ({rows_bad.numel()}) "
                        f"cols {cols_bad[:6].tolist()}..{int(cols_bad[-1])} ({cols_bad.numel()}) | equals own partial {int(own.sum())}, peer partial {int(other.sum())}, "
                        f"zero {int(((got == 0) & bad).sum())}")
            print(msg, flush=True)
    comm.close()
if "timeline" in sys.argv:
    comm = peer.PeerComm(device=dev, max_m=2048 * world, max_n=4096, nvls=True, p2p=False)
    for (M, N, K) in [(2048, 4096, 4096), (2048 * world, 4096, 4096)]:
        Kl = K // world // 16 * 16
        a = torch.randint(-128, 128, (M, Kl), dtype=torch.int8, device=dev)
        w = torch.randint(-127, 128, (N, Kl), dtype=torch.int8, device=dev)
        dbg = torch.zeros(161 * 8, dtype=torch.int64, device=dev)
        for _ in range(5):
            comm.linear_q8_allreduce_nvls(a, w, None, 1e-4)
        torch.cuda.synchronize()
        dist.barrier()
        os.environ["ASQ_DEBUG_TIMELINE"] = hex(dbg.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        comm.linear_q8_allreduce_nvls(a, w, None, 1e-4)
        e1.record()
        torch.cuda.synchronize()
        del os.environ["ASQ_DEBUG_TIMELINE"]
        t = dbg.view(161, 8).cpu().double()
        used = t[:, 7] > 0          # real CTAs stamp slot 7; the row after them holds the handshake stamps
        n_used = int(used.sum())
        hs = t[n_used]
        t = t[used]
        t0 = t[:, 0].min()
        R = int(os.environ.get("ASQ_NVLS_REDUCERS", "32"))
        red = torch.zeros(n_used, dtype=torch.bool)
        red[n_used - R:] = True  # the reducer CTAs are the last R of the grid
        out = [f"rank {rank} {M}x{N}x{K}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us, CTAs {n_used}, reducers {int(red.sum())}"]
        def stat(name, col, rel=True):
            col = col[col > 0]
            if len(col):
                col = (col - t0) if rel else col
                out.append(f"   {name:34s} min {col.min() / 1e3:8.2f}  median {col.median() / 1e3:8.2f}  max {col.max() / 1e3:8.2f} us (n={len(col)})")
        g, r = t[~red], t[red]
        stat("gemm: start", g[:, 0]); stat("gemm: setup done", g[:, 1]); stat("gemm: first acc ready", g[:, 5])
        stat("gemm: time signalling (sum/CTA)", g[:, 2], rel=False); stat("gemm: last tile stored", g[:, 6]); stat("gemm: exit", g[:, 7])
        stat("reducer: first slab landed", r[:, 2]); stat("reducer: first slab reduced", r[:, 3])
        stat("reducer: waiting (sum, warp 0)", r[:, 4], rel=False); stat("reducer: reducing (sum, warp 0)", r[:, 5], rel=False)
        stat("reducer: done", r[:, 6]); stat("reducer: exit", r[:, 7]); stat("reducer: sys fence done", r[:, 1])
        out.append(f"   handshake (last CTA): entered {(hs[0] - t0) / 1e3:.2f}, peers told {(hs[1] - t0) / 1e3:.2f}, peers heard {(hs[2] - t0) / 1e3:.2f}, done {(hs[3] - t0) / 1e3:.2f} us")
        print("\n".join(out), flush=True)
    comm.close()
dist.barrier()
dist.destroy_process_group()
