"""
A small Kademlia DHT node (UDP, asyncio) — the control plane for out-of-box deployments.

The reference depends on the third-party ``kademlia`` package (/root/reference/lib/network/__init__.py:6,20), which is
neither vendored nor installable offline; this is an independent implementation of the protocol surface the library
needs (``listen``, ``bootstrap``, ``get``, ``set``): 160-bit SHA-1 ids, XOR metric, k-buckets (k=20, native C++ routing
table in csrc/host_runtime.cpp), iterative lookups with alpha=3, RPCs PING / STORE / FIND_NODE / FIND_VALUE.
Values carry a store timestamp; the newest one wins.  Messages are msgpack maps in single UDP datagrams — a
NON-EXECUTABLE encoding: a datagram from any host on the network is parsed, never unpickled (stored VALUES stay opaque
bytes; what they contain is the business of the layer above).  Every field of an incoming message is type- and
length-checked before it reaches the routing table (the native table copies exactly 20 id bytes).
"""
import asyncio
import ctypes
import hashlib
import os
import msgpack
import socket
import struct
import threading
import time
from typing import Dict, List, Optional, Sequence, Tuple

from ..ops import host

K_BUCKET = 20
ALPHA = 3
RPC_TIMEOUT = 1.0
VALUE_TTL = 3600.0


MAX_VALUE_BYTES = 60000   # one UDP datagram


def _is_id(x) -> bool:
    return isinstance(x, bytes) and len(x) == 20


def _valid_message(msg: dict) -> bool:
    """types / lengths of every field of an incoming RPC (untrusted input)"""
    if msg.get("t") not in ("q", "r") or not isinstance(msg.get("id"), bytes) or len(msg["id"]) != 8:
        return False
    if not _is_id(msg.get("sender")):
        return False
    for field in ("key", "target"):
        if field in msg and not _is_id(msg[field]):
            return False
    if "value" in msg and not isinstance(msg["value"], bytes):
        return False
    if "ts" in msg and not isinstance(msg["ts"], (int, float)):
        return False
    if msg["t"] == "q":
        m = msg.get("m")
        if m not in ("ping", "store", "find_node", "find_value"):
            return False
        if m == "store" and not ("key" in msg and "value" in msg and "ts" in msg):
            return False
        if m == "find_node" and "target" not in msg:
            return False
        if m == "find_value" and "key" not in msg:
            return False
    if "nodes" in msg:
        nodes = msg["nodes"]
        if not isinstance(nodes, list):
            return False
        for entry in nodes:
            if not (isinstance(entry, list) and len(entry) == 2 and _is_id(entry[0]) and isinstance(entry[1], list)
                    and len(entry[1]) == 2 and isinstance(entry[1][0], str) and isinstance(entry[1][1], int)
                    and 0 < entry[1][1] < 65536):
                return False
            try:
                socket.inet_aton(entry[1][0])
            except OSError:
                return False
    return True


def sha1(data) -> bytes:
    if isinstance(data, str):
        data = data.encode()
    return hashlib.sha1(data).digest()


def xor_distance(a: bytes, b: bytes) -> int:
    return int.from_bytes(a, "big") ^ int.from_bytes(b, "big")


def _ip_to_int(ip: str) -> int:
    return struct.unpack("!I", socket.inet_aton(ip))[0]


def _int_to_ip(n: int) -> str:
    return socket.inet_ntoa(struct.pack("!I", n))


class RoutingTable:
    """k-bucket routing table; native when liblah_host.so is available, otherwise a Python fallback"""

    def __init__(self, node_id: bytes, k: int = K_BUCKET):
        self.node_id, self.k = node_id, k
        self._lib = host.lib()
        self._handle = self._lib.lah_rt_create(node_id, k) if self._lib is not None else None
        self._py: Dict[bytes, Tuple[str, int, float]] = {}

    def add(self, node_id: bytes, addr: Tuple[str, int]) -> bool:
        if not _is_id(node_id) or node_id == self.node_id:
            return False
        if self._handle is not None:
            evict_id = ctypes.create_string_buffer(20)
            evict_ip, evict_port = ctypes.c_uint32(), ctypes.c_uint16()
            r = self._lib.lah_rt_add(self._handle, node_id, _ip_to_int(addr[0]), addr[1], time.time(), evict_id,
                                     ctypes.byref(evict_ip), ctypes.byref(evict_port))
            return r == 1
        self._py[node_id] = (addr[0], addr[1], time.time())
        return True

    def remove(self, node_id: bytes):
        if self._handle is not None:
            self._lib.lah_rt_remove(self._handle, node_id)
        else:
            self._py.pop(node_id, None)

    def closest(self, target: bytes, n: int = K_BUCKET) -> List[Tuple[bytes, Tuple[str, int]]]:
        if not _is_id(target):
            raise ValueError("routing-table targets are 20-byte ids")
        if self._handle is not None:
            ids = ctypes.create_string_buffer(20 * n)
            ips, ports = (ctypes.c_uint32 * n)(), (ctypes.c_uint16 * n)()
            m = self._lib.lah_rt_closest(self._handle, target, n, ids, ips, ports)
            return [(ids.raw[20 * i: 20 * i + 20], (_int_to_ip(ips[i]), int(ports[i]))) for i in range(m)]
        items = sorted(self._py.items(), key=lambda kv: xor_distance(kv[0], target))[:n]
        return [(nid, (ip, port)) for nid, (ip, port, _) in items]

    def __len__(self):
        return self._lib.lah_rt_size(self._handle) if self._handle is not None else len(self._py)

    def __del__(self):
        if getattr(self, "_handle", None) is not None and self._lib is not None:
            self._lib.lah_rt_destroy(self._handle)
            self._handle = None


class _Protocol(asyncio.DatagramProtocol):
    def __init__(self, node: "DHTNode"):
        self.node = node
        self.transport = None

    def connection_made(self, transport):
        self.transport = transport

    def datagram_received(self, data, addr):
        try:
            msg = msgpack.unpackb(data, raw=False, strict_map_key=True, max_bin_len=MAX_VALUE_BYTES, max_str_len=64,
                                  max_array_len=4 * K_BUCKET, max_map_len=16)
        except Exception:
            return
        if not isinstance(msg, dict) or not _valid_message(msg):
            return
        try:
            self.node._on_message(msg, addr)
        except Exception:   # a malformed but well-typed message must not take the endpoint down
            return


class DHTNode:
    """async API mirroring kademlia.network.Server: listen / bootstrap / get / set"""

    def __init__(self, node_id: Optional[bytes] = None):
        self.node_id = node_id or sha1(os.urandom(32))
        self.table = RoutingTable(self.node_id)
        self.storage: Dict[bytes, Tuple[bytes, float]] = {}
        self._pending: Dict[bytes, asyncio.Future] = {}
        self._protocol: Optional[_Protocol] = None
        self.port = None

    # ------------------------------------------------------------------ transport
    async def listen(self, port: int, interface: str = "0.0.0.0"):
        loop = asyncio.get_event_loop()
        _, self._protocol = await loop.create_datagram_endpoint(lambda: _Protocol(self), local_addr=(interface, port))
        self.port = self._protocol.transport.get_extra_info("sockname")[1]

    def stop(self):
        if self._protocol is not None and self._protocol.transport is not None:
            self._protocol.transport.close()

    def _send(self, msg: dict, addr):
        self._protocol.transport.sendto(msgpack.packb(msg, use_bin_type=True), addr)

    async def _rpc(self, addr, method: str, **payload) -> Optional[dict]:
        rpc_id = os.urandom(8)
        future = asyncio.get_event_loop().create_future()
        self._pending[rpc_id] = future
        self._send(dict(t="q", id=rpc_id, m=method, sender=self.node_id, **payload), addr)
        try:
            return await asyncio.wait_for(future, RPC_TIMEOUT)
        except asyncio.TimeoutError:
            return None
        finally:
            self._pending.pop(rpc_id, None)

    def _on_message(self, msg: dict, addr):
        sender = msg.get("sender")
        if _is_id(sender):
            self.table.add(sender, (addr[0], addr[1]))
        if msg.get("t") == "r":
            future = self._pending.get(msg.get("id"))
            if future is not None and not future.done():
                future.set_result(msg)
            return
        reply = dict(t="r", id=msg.get("id"), sender=self.node_id)
        method = msg.get("m")
        if method == "ping":
            pass
        elif method == "store":
            self._store_local(msg["key"], msg["value"], msg["ts"])
        elif method == "find_node":
            reply["nodes"] = [[nid, [ip, port]] for nid, (ip, port) in self.table.closest(msg["target"], K_BUCKET)]
        elif method == "find_value":
            hit = self.storage.get(msg["key"])
            if hit is not None and time.time() - hit[1] <= VALUE_TTL:
                reply["value"], reply["ts"] = hit
            else:
                reply["nodes"] = [[nid, [ip, port]] for nid, (ip, port) in self.table.closest(msg["key"], K_BUCKET)]
        else:
            return
        self._send(reply, addr)

    def _store_local(self, key: bytes, value: bytes, ts: float):
        old = self.storage.get(key)
        if old is None or old[1] <= ts:
            self.storage[key] = (value, ts)

    # ------------------------------------------------------------------ kademlia operations
    async def bootstrap(self, peers: Sequence[Tuple[str, int]]):
        for host_, port in peers:
            ip = socket.gethostbyname(host_)
            await self._rpc((ip, port), "ping")
        if peers:
            await self._lookup(self.node_id, find_value=False)

    async def _lookup(self, target: bytes, find_value: bool):
        """iterative node/value lookup; returns (value_or_None, closest_contacts)"""
        shortlist = {nid: addr for nid, addr in self.table.closest(target, K_BUCKET)}
        queried = set()
        best_value = None
        while True:
            candidates = sorted((nid for nid in shortlist if nid not in queried),
                                key=lambda nid: xor_distance(nid, target))[:ALPHA]
            if not candidates:
                break
            queried.update(candidates)
            method = "find_value" if find_value else "find_node"
            payload = dict(key=target) if find_value else dict(target=target)
            replies = await asyncio.gather(*(self._rpc(shortlist[nid], method, **payload) for nid in candidates))
            for nid, reply in zip(candidates, replies):
                if reply is None:
                    self.table.remove(nid)
                    shortlist.pop(nid, None)
                    continue
                if "value" in reply and (best_value is None or reply["ts"] > best_value[1]):
                    best_value = (reply["value"], reply["ts"])
                for other_id, other_addr in reply.get("nodes", ()):
                    if other_id != self.node_id and other_id not in shortlist:
                        shortlist[other_id] = tuple(other_addr)
            if best_value is not None:
                break
            closest = sorted(shortlist, key=lambda nid: xor_distance(nid, target))[:K_BUCKET]
            if all(nid in queried for nid in closest):
                break
        closest = sorted(shortlist.items(), key=lambda kv: xor_distance(kv[0], target))[:K_BUCKET]
        return best_value, closest

    async def set(self, key, value: bytes) -> bool:
        digest, ts = sha1(key), time.time()
        _, closest = await self._lookup(digest, find_value=False)
        # store locally when we are among the k closest (or alone), then on the closest peers
        if len(closest) < K_BUCKET or xor_distance(self.node_id, digest) < xor_distance(closest[-1][0], digest):
            self._store_local(digest, value, ts)
        if closest:
            await asyncio.gather(*(self._rpc(addr, "store", key=digest, value=value, ts=ts) for _, addr in closest))
        return True

    async def get(self, key) -> Optional[bytes]:
        digest = sha1(key)
        local = self.storage.get(digest)
        found, _ = await self._lookup(digest, find_value=True)
        candidates = [c for c in (local, found) if c is not None and time.time() - c[1] <= VALUE_TTL]
        if not candidates:
            return None
        return max(candidates, key=lambda c: c[1])[0]
