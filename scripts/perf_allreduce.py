"""Fused GEMM+all-reduce vs GEMM + NCCL all-reduce, row-parallel shapes of Llama-2-7B at TP = world.
torchrun --nproc-per-node 2 scripts/perf_allreduce.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L  # noqa: E402
from autosmoothquant_b200 import peer  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)


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


Mmax = 2048 * world
nvls_only = "nvls-only" in sys.argv
comm = peer.PeerComm(device=dev, max_m=Mmax, max_n=4096)
comm_mc = peer.PeerComm(device=dev, max_m=Mmax, max_n=4096, multicast=True)
comm_nv = peer.PeerComm(device=dev, max_m=Mmax, max_n=4096, nvls=True, p2p=False)
for (M, N, K) in [(2048, 4096, 4096), (2048, 4096, 11008), (Mmax, 4096, 4096), (Mmax, 4096, 11008)]:
    Kl = K // world // 16 * 16
    a = torch.randint(-128, 128, (M, Kl), dtype=torch.int8, device=dev)
    w = torch.randint(-127, 128, (N, Kl), dtype=torch.int8, device=dev)

    def nccl():
        y = L.w8a8_linear_q8(a, w, None, 1e-4)
        dist.all_reduce(y)
        return y

    if nvls_only:
        t_nv = timeit(lambda: comm_nv.linear_q8_allreduce_nvls(a, w, None, 1e-4))
        if rank == 0:
            print(f"world {world}  {M}x{N}x{K} (K/rank {Kl}): fused NVLS (in-switch) {t_nv:6.1f} us  [ASQ_NVLS_REDUCERS={os.environ.get('ASQ_NVLS_REDUCERS', 'default')}]", flush=True)
        continue
    t_gemm = timeit(lambda: L.w8a8_linear_q8(a, w, None, 1e-4))
    y = L.w8a8_linear_q8(a, w, None, 1e-4)
    t_ar = timeit(lambda: dist.all_reduce(y))
    t_nccl = timeit(nccl)
    t_fused = timeit(lambda: comm.linear_q8_allreduce(a, w, None, 1e-4))
    t_fused16 = timeit(lambda: comm.linear_q8_allreduce(a, w, None, 1e-4, partials="native"))
    t_mc = timeit(lambda: comm_mc.linear_q8_allreduce(a, w, None, 1e-4))
    t_mc16 = timeit(lambda: comm_mc.linear_q8_allreduce(a, w, None, 1e-4, partials="native"))
    t_nv = timeit(lambda: comm_nv.linear_q8_allreduce_nvls(a, w, None, 1e-4))
    if rank == 0:
        print(f"world {world}  {M}x{N}x{K} (K/rank {Kl}): GEMM {t_gemm:6.1f} us | NCCL all-reduce {t_ar:6.1f} us | "
              f"GEMM+NCCL {t_nccl:6.1f} us | fused int32 p2p {t_fused:6.1f} | fused 16-bit p2p {t_fused16:6.1f} | "
              f"fused int32 multicast {t_mc:6.1f} | fused 16-bit multicast {t_mc16:6.1f} | fused NVLS (in-switch) {t_nv:6.1f} us", flush=True)
comm_nv.close()
comm_mc.close()
comm.close()
dist.destroy_process_group()
