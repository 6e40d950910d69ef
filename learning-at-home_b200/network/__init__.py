"""
TesseractNetwork — expert discovery: ``expert::<uid> -> ((host, port), timestamp)`` and ``prefix::<uid-prefix> ->
timestamp`` heartbeats (API parity: /root/reference/lib/network/__init__.py:11-129).

Two backends behind the same API:
  * DHT mode (default): a Kademlia node (network/dht.py) running on a background asyncio thread; unlike the reference
    (a forked process serving commands serially over an unlocked pipe) every public method is thread-safe and lookups of
    one call run concurrently;
  * in-box mode (``InBoxNetwork``): the "DHT collapse" of SURVEY.md §5.8 — a native hash index shared by all ranks of one
    box plus the device-resident liveness table read by the gate kernel (parallel/engine.py ``EngineContext.alive``).
"""
import asyncio
import datetime
import threading
import time
from concurrent.futures import Future
from typing import Dict, List, Optional, Sequence, Tuple

from ..utils import PickleSerializer
import io
import pickle


class _PlainUnpickler(pickle.Unpickler):
    """DHT values come from arbitrary peers: accept the reference's pickled ((host, port), datetime) / datetime records
    (lib/network/__init__.py:71-83 of the reference) and refuse every pickle that names anything but datetime.datetime"""

    def find_class(self, module, name):
        if (module, name) == ("datetime", "datetime"):
            return datetime.datetime
        raise pickle.UnpicklingError(f"DHT values may not reference {module}.{name}")


def _loads_value(raw: bytes):
    return _PlainUnpickler(io.BytesIO(raw)).load()
from .dht import DHTNode

UID_DELIMETER = "."
HEARTBEAT_EXPIRATION = 120


class TesseractNetwork:
    UID_DELIMETER = UID_DELIMETER  # splits expert uids over this delimeter
    HEARTBEAT_EXPIRATION = HEARTBEAT_EXPIRATION  # an expert is inactive iff it posted no heartbeat for this many seconds
    make_key = "{}::{}".format

    def __init__(self, *initial_peers: Tuple[str, int], port=8081, start=False):
        self.port, self.initial_peers = port, initial_peers
        self.node = DHTNode()
        self._loop: Optional[asyncio.AbstractEventLoop] = None
        self._thread: Optional[threading.Thread] = None
        self._ready = threading.Event()
        if start:
            self.start()

    # ------------------------------------------------------------------ lifecycle (process-like API)
    def start(self):
        if self.is_alive():
            return
        self._thread = threading.Thread(target=self.run, name=f"TesseractNetwork:{self.port}", daemon=True)
        self._thread.start()
        if not self._ready.wait(timeout=30):
            raise RuntimeError("DHT node failed to start")

    def run(self) -> None:
        loop = asyncio.new_event_loop()
        asyncio.set_event_loop(loop)
        self._loop = loop
        loop.run_until_complete(self.node.listen(self.port))
        self.port = self.node.port
        try:
            loop.run_until_complete(self.node.bootstrap(self.initial_peers))
        finally:
            self._ready.set()
        loop.run_forever()
        self.node.stop()
        loop.close()

    def is_alive(self) -> bool:
        return self._thread is not None and self._thread.is_alive()

    def shutdown(self):
        if self._loop is not None and self.is_alive():
            self._loop.call_soon_threadsafe(self._loop.stop)
            self._thread.join(timeout=5)

    def join(self, timeout=None):
        if self._thread is not None:
            self._thread.join(timeout)

    def _submit(self, coroutine) -> Future:
        assert self.is_alive(), "TesseractNetwork is not running; call start() first"
        return asyncio.run_coroutine_threadsafe(coroutine, self._loop)

    # ------------------------------------------------------------------ experts
    def get_experts(self, uids: List[str], heartbeat_expiration=HEARTBEAT_EXPIRATION) -> List[Optional["RemoteExpert"]]:
        """Find experts by uid; returns [RemoteExpert or None (unknown / heartbeat expired)]"""
        return self._submit(self._get_experts(list(uids), heartbeat_expiration)).result()

    async def _get_experts(self, uids, heartbeat_expiration):
        from ..client.remote_expert import RemoteExpert
        found = await asyncio.gather(*(self.node.get(self.make_key("expert", uid)) for uid in uids))
        now = datetime.datetime.now()
        experts = []
        for uid, raw in zip(uids, found):
            expert = None
            if raw is not None:
                (host, port), timestamp = _loads_value(raw)
                if (now - timestamp).total_seconds() <= heartbeat_expiration:
                    expert = RemoteExpert(uid=uid, host=host, port=port)
            experts.append(expert)
        return experts

    def declare_experts(self, uids: Sequence[str], addr, port, wait_timeout=0):
        """Publish (or refresh the heartbeat of) experts and of all their uid prefixes.  Fire-and-forget unless
        wait_timeout > 0."""
        future = self._submit(self._declare_experts(list(uids), addr, port))
        if wait_timeout:
            future.result(timeout=wait_timeout)

    async def _declare_experts(self, uids, addr, port):
        timestamp = datetime.datetime.now()
        expert_meta = PickleSerializer.dumps(((addr, port), timestamp))
        prefix_meta = PickleSerializer.dumps(timestamp)
        prefixes = set()
        for uid in uids:
            parts = uid.split(self.UID_DELIMETER)
            prefixes.update(self.UID_DELIMETER.join(parts[:i + 1]) for i in range(len(parts)))
        await asyncio.gather(*(self.node.set(self.make_key("expert", uid), expert_meta) for uid in uids),
                             *(self.node.set(self.make_key("prefix", prefix), prefix_meta) for prefix in prefixes))

    def first_k_active(self, prefixes: List[str], k: int, heartbeat_expiration=HEARTBEAT_EXPIRATION, max_prefetch=None):
        """The first k prefixes (in the given priority order) that have a fresh heartbeat; used by the beam search."""
        return self._submit(self._first_k_active(list(prefixes), k, heartbeat_expiration, max_prefetch or k)).result()

    async def _first_k_active(self, prefixes, k, heartbeat_expiration, max_prefetch):
        lookups: Dict[int, asyncio.Task] = {}
        next_to_issue = 0

        def issue():
            nonlocal next_to_issue
            if next_to_issue < len(prefixes):
                lookups[next_to_issue] = asyncio.ensure_future(self.node.get(self.make_key("prefix", prefixes[next_to_issue])))
                next_to_issue += 1

        for _ in range(max(1, max_prefetch)):
            issue()
        active = []
        for i, prefix in enumerate(prefixes):
            raw = await lookups.pop(i)
            if raw is not None:
                timestamp = _loads_value(raw)
                if (datetime.datetime.now() - timestamp).total_seconds() <= heartbeat_expiration:
                    active.append(prefix)
                    if len(active) >= k:
                        break
            issue()  # keep the pipeline full
        for task in lookups.values():
            task.cancel()
        return active


from .inbox import InBoxNetwork  # noqa: E402

__all__ = ["TesseractNetwork", "InBoxNetwork", "UID_DELIMETER", "HEARTBEAT_EXPIRATION", "DHTNode"]
