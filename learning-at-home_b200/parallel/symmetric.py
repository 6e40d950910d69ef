"""
Symmetric heap: one device arena per rank, allocated at identical offsets on every rank and mapped into every peer
process with CUDA IPC, so that the kernels in csrc/moe.cu / csrc/adam.cu can load/store peer HBM directly over
NVLink 5 / NVSwitch.  ``torch.distributed`` (NCCL) is used only to exchange the 64-byte IPC handles and for barriers.

This replaces the reference's transport stack (TCP sockets + torch.save, /root/reference/lib/utils/connection.py and
lib/utils/serializer.py) and its shared-memory staging (lib/utils/shared_arrays.py) on the in-box fast path.

Allocation: on multi-GPU runs the arena comes from ``torch.distributed._symmetric_memory`` (CUDA VMM allocation +
rendezvous — used as ALLOCATOR ONLY), which also binds it to an NVSwitch MULTICAST object: ``mc_base`` is
the multicast alias of the same offsets, the address space of the ``multimem.*`` instructions in csrc/nvls.cu (in-switch
gradient reduction, one-store broadcast of flags / heartbeats).  If the VMM rendezvous is unavailable the arena falls back
to ``cudaMalloc`` + CUDA IPC (no multicast; the kernels then use unicast P2P loads / stores).
"""
import ctypes
import os
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

        self.mc_base = 0          # multicast alias of the arena (0: not available)
        self._symm = None         # (tensor, handle) when allocated through torch symmetric memory
        self._opened = []
        if self.world > 1 and not os.environ.get("LAH_HEAP_IPC"):
            try:
                self._alloc_vmm(group)
            except Exception as e:  # noqa: VMM / fabric handles unavailable -> legacy IPC arena
                if self.rank == 0:
                    print(f"[lah_b200] symmetric-memory rendezvous unavailable ({type(e).__name__}: {e}); using CUDA IPC",
                          flush=True)
                self._symm = None
        if self._symm is not None:
            kernels.set_peers(self.peer_bases, self.rank)
            kernels.set_multicast(self.mc_base)
            self._cursor = 0
            return

        base = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            native.check(lib.lah_symm_alloc(self.nbytes, ctypes.byref(base)), "lah_symm_alloc")
        self.base = int(base.value)
        self.peer_bases: List[int] = [self.base]

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
        kernels.set_multicast(0)
        self._bytes = torch.as_tensor(_CudaBuffer(self.base, self.nbytes), device=self.device)
        self._cursor = 0

    def _alloc_vmm(self, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        pg = group if group is not None else dist.group.WORLD
        t = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        handle = symm_mem.rendezvous(t, pg.group_name)
        t.zero_()
        torch.cuda.synchronize(self.device)
        self._symm = (t, handle)
        self._bytes = t
        self.base = int(handle.buffer_ptrs[self.rank])
        assert self.base == t.data_ptr()
        self.peer_bases = [int(p) for p in handle.buffer_ptrs]
        self.mc_base = int(handle.multicast_ptr or 0)

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
        if self._symm is not None:   # VMM arena: owned by torch's symmetric-memory allocator
            torch.cuda.synchronize(self.device)
            if self.world > 1:   # nobody unmaps while a peer may still be reading
                import torch.distributed as dist
                dist.barrier(group=self.group)
            self._bytes = None
            self._symm = None
            self.base = 0
            kernels.set_multicast(0)
            return
        if self.world > 1 and self.base:
            torch.cuda.synchronize(self.device)
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier(group=self.group)
        for p in self._opened:
            lib.lah_symm_close_handle(ctypes.c_void_p(p))
        self._opened = []
        if self.base:
            self._bytes = None
            lib.lah_symm_free(ctypes.c_void_p(self.base))
            self.base = 0
