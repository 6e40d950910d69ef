"""
Symmetric heap: one device arena per rank, allocated at identical offsets on every rank and mapped into every peer
process with CUDA IPC, so that the kernels in csrc/moe.cu / csrc/adam.cu can load/store peer HBM directly over
NVLink 5 / NVSwitch.  ``torch.distributed`` (NCCL) is used only to exchange the 64-byte IPC handles and for barriers.

This replaces the reference's transport stack (TCP sockets + torch.save, /root/reference/lib/utils/connection.py and
lib/utils/serializer.py) and its shared-memory staging (lib/utils/shared_arrays.py) on the in-box fast path.
"""
import ctypes
from typing import List, Optional, Tuple

import torch

from ..ops import kernels, native

_ALIGN = 1024


class _CudaBuffer:
    """exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without copying"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = dict(shape=(nbytes,), typestr="|u1", data=(ptr, False), version=2)


class SymmetricHeap:
    def __init__(self, nbytes: int, group=None, device: Optional[torch.device] = None):
        """
        :param nbytes: arena size per rank
        :param group: torch.distributed process group (None => single process, world size 1)
        """
        import torch.distributed as dist
        self.group = group
        distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if distributed else 1
        self.rank = dist.get_rank(group) if distributed else 0
        assert self.world <= kernels.MAX_WORLD, f"at most {kernels.MAX_WORLD} ranks per box"
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.nbytes = (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        lib = kernels._lib()

        base = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            native.check(lib.lah_symm_alloc(self.nbytes, ctypes.byref(base)), "lah_symm_alloc")
        self.base = int(base.value)
        self.peer_bases: List[int] = [self.base]
        self._opened = []

        if self.world > 1:
            handle = ctypes.create_string_buffer(64)
            native.check(lib.lah_symm_get_handle(ctypes.c_void_p(self.base), handle), "lah_symm_get_handle")
            mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=self.device)
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=group)
            self.peer_bases = []
            for r, h in enumerate(gathered):
                if r == self.rank:
                    self.peer_bases.append(self.base)
                    continue
                raw = bytes(h.cpu().tolist())
                p = ctypes.c_void_p()
                native.check(lib.lah_symm_open_handle(raw, ctypes.byref(p)), f"lah_symm_open_handle(rank {r})")
                self._opened.append(int(p.value))
                self.peer_bases.append(int(p.value))
        kernels.set_peers(self.peer_bases, self.rank)
        self._bytes = torch.as_tensor(_CudaBuffer(self.base, self.nbytes), device=self.device)
        self._cursor = 0

    # ------------------------------------------------------------------ allocation (identical order on all ranks!)
    def alloc(self, shape, dtype) -> Tuple[torch.Tensor, int]:
        """carve a tensor out of the arena; returns (tensor, byte offset inside the heap)"""
        numel = 1
        for s in shape:
            numel *= int(s)
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        off = (self._cursor + _ALIGN - 1) // _ALIGN * _ALIGN
        if off + nbytes > self.nbytes:
            raise MemoryError(f"symmetric heap exhausted: need {off + nbytes} of {self.nbytes} bytes")
        self._cursor = off + nbytes
        t = self._bytes[off: off + nbytes].view(dtype).view(*shape)
        return t, off

    def barrier(self):
        """host-side barrier (NCCL) + device sync; used at setup / teardown only"""
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)
        torch.cuda.synchronize(self.device)

    def close(self):
        lib = kernels._lib()
        for p in self._opened:
            lib.lah_symm_close_handle(ctypes.c_void_p(p))
        self._opened = []
        if self.base:
            self._bytes = None
            lib.lah_symm_free(ctypes.c_void_p(self.base))
            self.base = 0
