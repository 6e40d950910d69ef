"""ctypes bindings for the C++ host runtime (csrc/host_runtime.cpp) with pure-Python fallbacks."""
import ctypes
import os
from typing import List, Sequence

import torch

from . import native

_lib = None
_tried = False
c_ull = ctypes.c_ulonglong


def lib():
    """the native host library, or None when it cannot be built/loaded (then the Python fallbacks run)"""
    global _lib, _tried
    if _tried:
        return _lib
    _tried = True
    if os.environ.get("LAH_NO_NATIVE_HOST"):
        return None
    try:
        L = native.host_lib()
    except Exception:  # pragma: no cover - compiler missing
        return None
    V, I, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    L.lah_host_gather.restype = I
    L.lah_host_gather.argtypes = [V, V, I, V, I]
    L.lah_host_scatter.restype = I
    L.lah_host_scatter.argtypes = [V, V, V, I]
    L.lah_host_send_all.restype = ctypes.c_longlong
    L.lah_host_send_all.argtypes = [I, V, c_ull]
    L.lah_host_recv_exact.restype = ctypes.c_longlong
    L.lah_host_recv_exact.argtypes = [I, V, c_ull]
    L.lah_host_send_message.restype = I
    L.lah_host_send_message.argtypes = [I, ctypes.c_char_p, V, c_ull]
    L.lah_host_recv_header.restype = I
    L.lah_host_recv_header.argtypes = [I, ctypes.c_char_p, ctypes.POINTER(c_ull)]
    L.lah_rt_create.restype = V
    L.lah_rt_create.argtypes = [ctypes.c_char_p, I]
    L.lah_rt_destroy.restype = None
    L.lah_rt_destroy.argtypes = [V]
    L.lah_rt_add.restype = I
    L.lah_rt_add.argtypes = [V, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint16, D, V, V, V]
    L.lah_rt_remove.restype = I
    L.lah_rt_remove.argtypes = [V, ctypes.c_char_p]
    L.lah_rt_size.restype = I
    L.lah_rt_size.argtypes = [V]
    L.lah_rt_closest.restype = I
    L.lah_rt_closest.argtypes = [V, ctypes.c_char_p, I, V, V, V]
    L.lah_hash_bytes.restype = c_ull
    L.lah_hash_bytes.argtypes = [ctypes.c_char_p, I]
    L.lah_index_create.restype = V
    L.lah_index_create.argtypes = [I]
    L.lah_index_destroy.restype = None
    L.lah_index_destroy.argtypes = [V]
    L.lah_index_put.restype = I
    L.lah_index_put.argtypes = [V, c_ull, I, I, D]
    L.lah_index_get.restype = I
    L.lah_index_get.argtypes = [V, c_ull, D, D, ctypes.POINTER(I), ctypes.POINTER(I), ctypes.POINTER(D)]
    _lib = L
    return _lib


def gather_rows(parts: Sequence[torch.Tensor], out: torch.Tensor, threads: int = 4) -> torch.Tensor:
    """out[:sum(rows)] = cat(parts, dim=0) for CPU tensors; native multi-threaded memcpy when available"""
    L = lib()
    contiguous = all(p.device.type == "cpu" and p.is_contiguous() and p.dtype == out.dtype for p in parts)
    if L is None or not contiguous or not out.is_contiguous():
        torch.cat([p.to(out.dtype) for p in parts], dim=0, out=out)
        return out
    n = len(parts)
    srcs = (ctypes.c_void_p * n)(*[p.data_ptr() for p in parts])
    sizes = (c_ull * n)(*[p.numel() * p.element_size() for p in parts])
    assert sum(sizes) == out.numel() * out.element_size(), "gather_rows: size mismatch"
    L.lah_host_gather(ctypes.cast(srcs, ctypes.c_void_p), ctypes.cast(sizes, ctypes.c_void_p), n,
                      ctypes.c_void_p(out.data_ptr()), threads)
    return out
