"""
Raw-tensor wire format of the TCP fallback path — a negotiated EXTENSION of the reference protocol.

The reference serialises every request / reply with ``torch.save`` (/root/reference/lib/utils/serializer.py:34-41): for
the 8 MiB tensors of the throughput experiment that is four pickle passes per round trip (~12 ms each on the build
container).  A server of this package advertises ``tensor_wire=1`` in its 'info' reply; a ``RemoteExpert`` that sees the
flag sends 'fwdT' / 'bwdT' requests whose payload is this frame and gets 'resT' replies:

    frame   := magic 'LAHT' | u16 uid_len | uid | u32 n | n x tensor
    tensor  := u8 dtype_code | u8 ndim | ndim x i64 dims | u64 nbytes | raw bytes (C-contiguous, little endian)

Reference clients keep talking 'fwd_' / 'bwd_' + ``torch.save`` to the same server, and a RemoteExpert that meets a
reference server (no flag) falls back to it, so both directions stay wire compatible.  Tensors are sent straight from
their storage (one ``sendall`` per tensor, no concatenation) and received into one ``bytearray`` that the returned tensors
alias (``torch.frombuffer``): no copy on either side.
"""
import struct
from typing import List, Sequence, Tuple

import numpy as np
import torch

MAGIC = b"LAHT"
_DTYPES = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.int16, torch.int8,
           torch.uint8, torch.bool]
_CODE = {dt: i for i, dt in enumerate(_DTYPES)}
MAX_TENSORS, MAX_NDIM = 4096, 8
REQUEST_HEADERS = {"fwdT": "fwd_", "bwdT": "bwd_"}
REPLY_HEADER = "resT"


def supported(tensors: Sequence[torch.Tensor]) -> bool:
    return all(isinstance(t, torch.Tensor) and t.dtype in _CODE and not t.is_sparse for t in tensors)


def _as_bytes(t: torch.Tensor) -> memoryview:
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        t = t.view(torch.int16)
    elif t.dtype == torch.bool:
        t = t.view(torch.uint8)
    return memoryview(t.numpy()).cast("B") if t.numel() else memoryview(b"")


def encode(uid: str, tensors: Sequence[torch.Tensor]) -> Tuple[List[memoryview], int]:
    """-> (buffers to send back to back, total length)"""
    uid_b = uid.encode()
    parts: List = [MAGIC + struct.pack("<H", len(uid_b)) + uid_b + struct.pack("<I", len(tensors))]
    for t in tensors:
        data = _as_bytes(t)
        parts.append(struct.pack(f"<BB{t.dim()}qQ", _CODE[t.dtype], t.dim(), *t.shape, data.nbytes))
        if data.nbytes:
            parts.append(data)
    return parts, sum(len(p) if isinstance(p, bytes) else p.nbytes for p in parts)


def decode(buf) -> Tuple[str, Tuple[torch.Tensor, ...]]:
    """``buf``: bytes-like (kept alive by the returned tensors, which alias it)"""
    view = memoryview(buf)
    total = view.nbytes

    def need(off, n):   # every field is attacker-controlled: check it against the bytes actually received
        if n < 0 or off + n > total:
            raise ValueError("truncated or inconsistent tensor frame")

    need(0, 6)
    if bytes(view[:4]) != MAGIC:
        raise ValueError("not a tensor frame")
    (uid_len,) = struct.unpack_from("<H", view, 4)
    off = 6
    need(off, uid_len + 4)
    uid = bytes(view[off: off + uid_len]).decode()
    off += uid_len
    (n,) = struct.unpack_from("<I", view, off)
    off += 4
    if n > MAX_TENSORS:
        raise ValueError(f"frame announces {n} tensors (limit {MAX_TENSORS})")
    out = []
    for _ in range(n):
        need(off, 2)
        code, ndim = struct.unpack_from("<BB", view, off)
        off += 2
        if code >= len(_DTYPES) or ndim > MAX_NDIM:
            raise ValueError("bad dtype code / rank in tensor frame")
        need(off, 8 * ndim + 8)
        dims = struct.unpack_from(f"<{ndim}q", view, off)
        off += 8 * ndim
        (nbytes,) = struct.unpack_from("<Q", view, off)
        off += 8
        dtype = _DTYPES[code]
        numel = 1
        for d in dims:
            if d < 0:
                raise ValueError("negative dimension in tensor frame")
            numel *= d
        if numel * torch.empty((), dtype=dtype).element_size() != nbytes:   # dims must describe exactly the bytes sent
            raise ValueError("tensor dims do not match its byte count")
        need(off, nbytes)
        if nbytes:
            carrier = torch.int16 if dtype == torch.bfloat16 else torch.uint8 if dtype == torch.bool else dtype
            t = torch.frombuffer(view[off: off + nbytes], dtype=carrier)
            if carrier != dtype:
                t = t.view(dtype)
            t = t.view(tuple(dims))      # tuple form: a 0-dim tensor has dims == ()
        else:
            t = torch.empty(tuple(dims), dtype=dtype)
        off += nbytes
        out.append(t)
    return uid, tuple(out)
