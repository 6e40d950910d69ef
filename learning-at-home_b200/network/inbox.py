"""
InBoxNetwork — the in-box replacement of the Kademlia DHT (SURVEY.md §5.8 "DHT collapse").

Inside one NVSwitch box every expert is a known (owner rank, local slot); discovery is an O(1) lookup in a native hash
index (csrc/host_runtime.cpp ``lah_index_*``) and liveness is a heartbeat timestamp per expert and per uid prefix.
The class implements the SAME methods as ``TesseractNetwork`` (``declare_experts`` / ``get_experts`` /
``first_k_active``), so ``GatingFunction`` and ``TesseractServer`` run unchanged on top of it; additionally
``alive_mask(grid, prefix)`` exports the liveness table that the fused gate kernel reads (``EngineContext.alive``).
"""
import ctypes
import threading
import time
from typing import Dict, List, Optional, Sequence, Tuple

from ..ops import host

UID_DELIMETER = "."
HEARTBEAT_EXPIRATION = 120


class InBoxNetwork:
    UID_DELIMETER = UID_DELIMETER
    HEARTBEAT_EXPIRATION = HEARTBEAT_EXPIRATION
    make_key = "{}::{}".format

    def __init__(self, *initial_peers, port=None, start=True):
        self._lib = host.lib()
        self._index = self._lib.lah_index_create(1024) if self._lib is not None else None
        self._py: Dict[int, Tuple[int, int, float]] = {}
        self._endpoints: Dict[str, Tuple[str, int]] = {}
        self._lock = threading.Lock()
        self._alive = bool(start)

    # process-like API of TesseractNetwork
    def start(self):
        self._alive = True

    def is_alive(self) -> bool:
        return self._alive

    def shutdown(self):
        self._alive = False

    def join(self, timeout=None):
        pass

    # ------------------------------------------------------------------ index primitives
    def _hash(self, kind: str, name: str) -> int:
        key = self.make_key(kind, name).encode()
        if self._lib is not None:
            return int(self._lib.lah_hash_bytes(key, len(key)))
        return hash(key) & (2 ** 63 - 1) or 1

    def _put(self, kind, name, owner=0, slot=0, now=None):
        now = time.time() if now is None else now
        h = self._hash(kind, name)
        if self._index is not None:
            self._lib.lah_index_put(self._index, h, owner, slot, now)
        else:
            with self._lock:
                self._py[h] = (owner, slot, now)

    def _fresh(self, kind, name, max_age) -> Optional[Tuple[int, int]]:
        h = self._hash(kind, name)
        if self._index is not None:
            owner, slot, hb = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
            ok = self._lib.lah_index_get(self._index, h, time.time(), float(max_age), ctypes.byref(owner),
                                         ctypes.byref(slot), ctypes.byref(hb))
            return (owner.value, slot.value) if ok else None
        with self._lock:
            entry = self._py.get(h)
        if entry is None or time.time() - entry[2] > max_age:
            return None
        return entry[0], entry[1]

    # ------------------------------------------------------------------ TesseractNetwork API
    def declare_experts(self, uids: Sequence[str], addr, port, wait_timeout=0, owner: int = 0, slots=None):
        now = time.time()
        for i, uid in enumerate(uids):
            self._put("expert", uid, owner, slots[i] if slots is not None else i, now)
            with self._lock:
                self._endpoints[uid] = (addr, port)
            parts = uid.split(self.UID_DELIMETER)
            for j in range(len(parts)):
                self._put("prefix", self.UID_DELIMETER.join(parts[:j + 1]), owner, 0, now)

    def get_experts(self, uids: List[str], heartbeat_expiration=HEARTBEAT_EXPIRATION):
        from ..client.remote_expert import RemoteExpert
        out = []
        for uid in uids:
            hit = self._fresh("expert", uid, heartbeat_expiration)
            if hit is None:
                out.append(None)
            else:
                host_, port = self._endpoints.get(uid, ("127.0.0.1", 0))
                out.append(RemoteExpert(uid=uid, host=host_, port=port))
        return out

    def first_k_active(self, prefixes: List[str], k: int, heartbeat_expiration=HEARTBEAT_EXPIRATION, max_prefetch=None):
        active = []
        for prefix in prefixes:
            if self._fresh("prefix", prefix, heartbeat_expiration) is not None:
                active.append(prefix)
                if len(active) >= k:
                    break
        return active

    # ------------------------------------------------------------------ export for the gate kernel
    def alive_mask(self, grid_size: Sequence[int], uid_prefix: str, heartbeat_expiration=HEARTBEAT_EXPIRATION):
        """uint8 tensor [prod(grid)] (row-major expert index): 1 where the expert's heartbeat is fresh"""
        import itertools
        import torch
        flags = []
        for coords in itertools.product(*(range(g) for g in grid_size)):
            uid = self.UID_DELIMETER.join([uid_prefix] + [str(c) for c in coords])
            flags.append(1 if self._fresh("expert", uid, heartbeat_expiration) is not None else 0)
        return torch.tensor(flags, dtype=torch.uint8)

    def __del__(self):
        if getattr(self, "_index", None) is not None and self._lib is not None:
            self._lib.lah_index_destroy(self._index)
            self._index = None
