"""Peer-memory communicator for the row-parallel linear fused with its all-reduce.

One process per GPU (``torch.distributed`` only carries the 64-byte CUDA IPC handles at set-up time): every
rank allocates a receive buffer, control words and two output buffers with ``asq_dev_alloc`` (plain
``cudaMalloc``, zero-filled), exports them, and maps the peers' buffers into its own address space.  After
that ``PeerComm.linear_q8_allreduce`` is ONE kernel launch per rank and no NCCL call: partial int32
accumulators travel by P2P stores over NVLink, finished tiles are TMA-stored into every rank's output
(``include/asq.h``: ``asq_w8a8_linear_q8_allreduce``).

``nvls=True`` adds the in-switch variant (``linear_q8_allreduce_nvls``): the partial products stay in this rank's
slice of a torch symmetric-memory allocation, the owner of a tile sums all ranks' copies with
``multimem.ld_reduce`` (the NVSwitch adds them) and broadcasts the result with ``multimem.st``.

Output buffers alternate between consecutive launches; a result stays valid until the launch after the next
one on the same communicator (the protocol's end-of-launch handshake guarantees every rank has consumed the
previous use by then, provided consumers are enqueued before the next launch — true for a sequential stream).
Callers that keep a result longer must copy it (``tp.RowParallelLinear`` does unless told otherwise).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib


class _DevBuffer:
    """cudaMalloc'ed memory exposed to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerComm:
    def __init__(self, group=None, device: Optional[torch.device] = None, max_m: int = 2048, max_n: int = 8192,
                 dtype: torch.dtype = torch.bfloat16, multicast: bool = False, nvls: bool = False, p2p: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("PeerComm needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if not 2 <= self.world <= 8:
            raise ValueError("PeerComm supports 2..8 ranks of one node")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_m, self.max_n, self.dtype = int(max_m), int(max_n), dtype
        lib = _lib.load()
        recv_b, ctl_b = ctypes.c_size_t(), ctypes.c_size_t()
        with torch.cuda.device(self.device):
            _lib._check(lib.asq_ar_buffer_bytes(self.max_m, self.max_n, self.world, ctypes.byref(recv_b), ctypes.byref(ctl_b)))
            y_bytes = self.max_m * self.max_n * 2
            # p2p=False (nvls only): no receive buffer for peer-stored partials is needed
            sizes = {"recv": recv_b.value if p2p else 1024, "ctl": ctl_b.value}
            self.p2p = p2p
            self._symm = None
            self.multicast_ptr = 0
            self._nvls = None
            if nvls:
                # [partials | y0 | y1] in ONE symmetric allocation: the partial buffer must be multicast-mapped for
                # multimem.ld_reduce, the outputs for multimem.st
                import torch.distributed._symmetric_memory as symm

                # + two banks of "slab landed" counters: 4 bytes per (64-row slab, 256-column tile), worst aspect ratio
                slabs = ((self.max_m + 63) // 64 + 3) * ((self.max_n + 255) // 256) + 4 * ((self.max_m * self.max_n) // (64 * 256) + 64)
                self._ctr_bank = (slabs * 4 + 1023) // 1024 * 1024
                self._nvls_tensor = symm.empty(3 * y_bytes + 2 * self._ctr_bank, dtype=torch.uint8, device=self.device)
                self._nvls_tensor.zero_()
                pg = dist.group.WORLD if group is None else group
                hdl = symm.rendezvous(self._nvls_tensor, pg.group_name)
                if not hdl.multicast_ptr:
                    raise RuntimeError("this system exposes no NVLS multicast address (nvls=True needs NVSwitch)")
                self._nvls = hdl
                self._nvls_mc = int(hdl.multicast_ptr)
                self._nvls_local = int(self._nvls_tensor.data_ptr())
                self._nvls_y_views = [self._nvls_tensor[y_bytes:2 * y_bytes], self._nvls_tensor[2 * y_bytes:3 * y_bytes]]
                self._nvls_launches = 0
            if multicast:
                # the output buffers live in torch symmetric memory so that an NVLS multicast address exists for them
                import torch.distributed._symmetric_memory as symm

                self._symm_tensor = symm.empty(2 * y_bytes, dtype=torch.uint8, device=self.device)
                self._symm_tensor.zero_()
                pg = dist.group.WORLD if group is None else group
                self._symm = symm.rendezvous(self._symm_tensor, pg.group_name)
                if not self._symm.multicast_ptr:
                    raise RuntimeError("this system exposes no NVLS multicast address (multicast=True needs NVSwitch)")
                self.multicast_ptr = int(self._symm.multicast_ptr)
            elif p2p:
                sizes.update({"y0": y_bytes, "y1": y_bytes})
            self._own = {}
            handles = {}
            for name, nbytes in sizes.items():
                ptr = ctypes.c_void_p()
                _lib._check(lib.asq_dev_alloc(nbytes, ctypes.byref(ptr)))
                self._own[name] = (ptr.value, nbytes)
                h = (ctypes.c_ubyte * 64)()
                _lib._check(lib.asq_ipc_export(ptr, h))
                handles[name] = bytes(h)
            torch.cuda.synchronize(self.device)
            gathered: List[dict] = [None] * self.world
            dist.all_gather_object(gathered, handles, group=group)
            self._opened = []
            self.ptrs = {name: [0] * self.world for name in sizes}
            for r in range(self.world):
                for name in sizes:
                    if r == self.rank:
                        self.ptrs[name][r] = self._own[name][0]
                    else:
                        out = ctypes.c_void_p()
                        buf = (ctypes.c_ubyte * 64).from_buffer_copy(gathered[r][name])
                        _lib._check(lib.asq_ipc_open(buf, ctypes.byref(out)))
                        self._opened.append(out.value)
                        self.ptrs[name][r] = out.value
        if self._symm is not None:
            base = [int(p) for p in self._symm.buffer_ptrs]
            self.ptrs["y0"] = base
            self.ptrs["y1"] = [b + y_bytes for b in base]
            self._y_views = [self._symm_tensor[:y_bytes], self._symm_tensor[y_bytes:]]
        elif p2p:
            self._y_views = [torch.as_tensor(_DevBuffer(*self._own[n]), device=self.device) for n in ("y0", "y1")]
        self._y_bytes = y_bytes
        self._tables = {name: (ctypes.c_void_p * self.world)(*self.ptrs[name]) for name in self.ptrs}
        self._launches = 0
        dist.barrier(group=group)  # every rank's buffers exist, are zeroed and mapped before the first launch

    def linear_q8_allreduce(self, xq: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                            dequant_scale: float, col_scale: Optional[torch.Tensor] = None,
                            row_scale: Optional[torch.Tensor] = None, partials: str = "int32") -> torch.Tensor:
        """sum over ranks of (xq_r . weight_r^T), dequantised (+ bias) — xq [M, K/world] int8, weight [N, K/world]
        int8.  partials="int32": exact integer exchange, bias on EVERY rank, result bit-identical to the unsharded
        module; partials="native": 16-bit dequantised partials (half the bytes, NCCL-native numerics), bias on ONE
        rank only.  Returns a [M, N] view of this rank's symmetric output buffer."""
        if partials not in ("int32", "native"):
            raise ValueError("partials must be 'int32' or 'native'")
        if not self.p2p:
            raise RuntimeError("this communicator was built with p2p=False (in-switch reduction only)")
        if xq.dtype != torch.int8 or weight.dtype != torch.int8 or xq.dim() != 2 or xq.shape[1] != weight.shape[1]:
            raise ValueError("linear_q8_allreduce expects int8 [M,K] activations and int8 [N,K] weights")
        if not (xq.is_cuda and weight.is_cuda and xq.is_contiguous() and weight.is_contiguous()):
            raise ValueError("linear_q8_allreduce expects contiguous CUDA tensors (no CPU fallback)")
        M, K = xq.shape
        N = weight.shape[0]
        if M > self.max_m or N > self.max_n or M * N > self.max_m * self.max_n:
            raise ValueError(f"[{M},{N}] exceeds the communicator's buffers [{self.max_m},{self.max_n}]")
        # the receive-buffer layout depends on (M, N): it must not exceed what was allocated for (max_m, max_n)
        lib = _lib.load()
        recv_b, ctl_b = ctypes.c_size_t(), ctypes.c_size_t()
        _lib._check(lib.asq_ar_buffer_bytes(M, N, self.world, ctypes.byref(recv_b), ctypes.byref(ctl_b)))
        if recv_b.value > self._own["recv"][1] or ctl_b.value > self._own["ctl"][1]:
            raise ValueError("shape needs larger exchange buffers than this communicator allocated")
        which = self._launches & 1
        y_table = self._tables["y1" if which else "y0"]
        with torch.cuda.device(self.device):
            rc = lib.asq_w8a8_linear_q8_allreduce(
                xq.data_ptr(), _lib._ptr(row_scale), weight.data_ptr(), _lib._ptr(bias), y_table, _lib._code(self.dtype),
                M, N, K, float(dequant_scale), _lib._ptr(col_scale), self._tables["recv"], self._tables["ctl"],
                self.rank, self.world, 1 if partials == "native" else 0,
                (self.multicast_ptr + which * self._y_bytes) if self.multicast_ptr else None, _lib._stream(self.device))
        _lib._check(rc)
        _lib._launches += 1
        self._launches += 1
        return self._y_views[which][: M * N * 2].view(self.dtype).view(M, N)

    def linear_q8_allreduce_nvls(self, xq: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                                 dequant_scale: float, col_scale: Optional[torch.Tensor] = None,
                                 row_scale: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """sum over ranks of T(dequant(xq_r . weight_r^T) (+ bias)) with the sum taken IN THE SWITCH — xq [M, K/world]
        int8 or float8_e4m3fn, weight [N, K/world] of the same 8-bit type; bias on ONE rank only.  One launch per
        rank: every rank's epilogue TMA-stores its 16-bit partial tiles into its own symmetric buffer and bumps a
        counter on the tile's owner (rank = tile % world); the owner's epilogue warps then multimem.ld_reduce the
        tile (fp32 accumulation in the NVSwitch) and multimem.st the sums into every rank's output.  The numerics
        are those of GEMM + ncclAllReduce(NVLS) in the 16-bit dtype.  Returns a [M, N] view of this rank's output
        buffer (alternating between two buffers)."""
        if self._nvls is None:
            raise RuntimeError("this communicator was built without nvls=True")
        fp8 = weight.dtype == torch.float8_e4m3fn
        if (xq.dtype != weight.dtype or xq.dtype not in (torch.int8, torch.float8_e4m3fn) or xq.dim() != 2
                or xq.shape[1] != weight.shape[1]):
            raise ValueError("linear_q8_allreduce_nvls expects 8-bit [M,K] activations and [N,K] weights of the same type")
        if not (xq.is_cuda and weight.is_cuda and xq.is_contiguous() and weight.is_contiguous()):
            raise ValueError("linear_q8_allreduce_nvls expects contiguous CUDA tensors (no CPU fallback)")
        out_dtype = out_dtype or self.dtype
        if out_dtype not in (torch.bfloat16, torch.float16):
            raise ValueError("linear_q8_allreduce_nvls reduces 16-bit partials (bf16 | f16)")
        M, K = xq.shape
        N = weight.shape[0]
        if M * N > self.max_m * self.max_n:
            raise ValueError(f"[{M},{N}] exceeds the communicator's buffers [{self.max_m},{self.max_n}]")
        lib = _lib.load()
        recv_b, ctl_b = ctypes.c_size_t(), ctypes.c_size_t()
        _lib._check(lib.asq_ar_buffer_bytes(M, N, self.world, ctypes.byref(recv_b), ctypes.byref(ctl_b)))
        if ctl_b.value > self._own["ctl"][1]:
            raise ValueError("shape needs more control words than this communicator allocated")
        which = self._nvls_launches & 1
        y_off = (1 + which) * self._y_bytes
        with torch.cuda.device(self.device):
            rc = lib.asq_q8_linear_allreduce_nvls(
                xq.data_ptr(), 1 if fp8 else 0, _lib._ptr(row_scale), weight.data_ptr(), _lib._ptr(bias),
                self._nvls_local, self._nvls_mc, self._nvls_mc + y_off, _lib._code(out_dtype), M, N, K, float(dequant_scale),
                _lib._ptr(col_scale), self._tables["ctl"], self._nvls_local + 3 * self._y_bytes, self._nvls_mc + 3 * self._y_bytes,
                self._ctr_bank, which, self.rank, self.world, _lib._stream(self.device))
        _lib._check(rc)
        _lib._launches += 1
        self._nvls_launches += 1
        return self._nvls_y_views[which][: M * N * 2].view(out_dtype).view(M, N)

    def close(self) -> None:
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for ptr in self._opened:
                lib.asq_ipc_close(ctypes.c_void_p(ptr))
            self._opened = []
            dist.barrier(group=self.group)
            for ptr, _ in self._own.values():
                lib.asq_dev_free(ctypes.c_void_p(ptr))
            self._own = {}
        # symmetric-memory allocations are released by torch once every reference is gone
        self._nvls = self._nvls_tensor = self._nvls_y_views = None
