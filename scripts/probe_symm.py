"""Probe torch symmetric memory / NVLS multicast availability: torchrun --nproc-per-node 2 scripts/probe_symm.py"""
import os
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.uint8, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs],
          "multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, flush=True)
except Exception as e:  # noqa: BLE001
    print(rank, "symmetric memory unavailable:", repr(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
