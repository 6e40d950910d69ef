"""
TesseractRuntime — the single loop that owns the device and executes batches formed by the experts' TaskPools
(API parity: /root/reference/lib/runtime/__init__.py:15-70).

Design differences from the reference:
  * no process zoo: pools are in-process queues, the scheduler is woken by a condition variable instead of a
    ``selectors`` loop over pipes, a prefetch thread assembles + uploads upcoming batches (pinned staging + async H2D on
    a side CUDA stream) while the main thread computes — the role of BackgroundGenerator(prefetch_batches) there;
  * scheduling is oldest-waiting-task-first across pools (the documented intent; the reference's max() picks the newest);
  * outputs are downloaded and delivered by ``sender_threads`` workers; runtime errors are delivered to the tasks'
    futures instead of leaving the clients hanging;
  * ``shutdown()`` makes ``main()`` return (the reference can only be interrupted).
"""
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from itertools import chain
from typing import Dict, Iterator, List, Optional, Tuple

import torch

from .expert_backend import ExpertBackend
from .task_pool import TaskPool, TaskPoolBase, BatchAssemblyError


class TesseractRuntime:
    def __init__(self, expert_backends: Dict[str, ExpertBackend], prefetch_batches: int = 64, sender_threads: int = 1,
                 device: torch.device = None):
        self.expert_backends = expert_backends
        self.pools: Tuple[TaskPool, ...] = tuple(chain(*(backend.get_pools() for backend in expert_backends.values())))
        self.device, self.prefetch_batches, self.sender_threads = device, prefetch_batches, sender_threads
        self._wakeup = threading.Condition()
        self._stop = threading.Event()
        self.ready = threading.Event()
        self.batches_processed = 0
        self.samples_processed = 0
        self.last_error: Optional[BaseException] = None

    # ------------------------------------------------------------------ scheduling
    def _notify(self, pool=None):
        with self._wakeup:
            self._wakeup.notify()

    def _next_pool(self, timeout: float = 0.05) -> Optional[TaskPool]:
        """the non-empty pool whose oldest task has waited longest, or None after :timeout: without work"""
        with self._wakeup:
            ready = [pool for pool in self.pools if not pool.empty]
            if not ready:
                self._wakeup.wait(timeout)
                ready = [pool for pool in self.pools if not pool.empty]
            return min(ready, key=lambda pool: pool.priority) if ready else None

    def iterate_minibatches_from_pools(self, timeout=None) -> Iterator[Tuple[TaskPool, int, List[torch.Tensor]]]:
        """yields (pool, batch_index, batch tensors on the device) in oldest-first order until shutdown"""
        copy_stream = torch.cuda.Stream(self.device) if self._on_cuda() else None
        while not self._stop.is_set():
            pool = self._next_pool()
            if pool is None:
                continue
            try:
                if copy_stream is not None:
                    with torch.cuda.stream(copy_stream):
                        batch_index, batch = pool.load_batch_to_runtime(timeout, self.device)
                        ready = torch.cuda.Event()
                        ready.record(copy_stream)
                    item = (pool, batch_index, batch, ready)
                else:
                    batch_index, batch = pool.load_batch_to_runtime(timeout, self.device)
                    item = (pool, batch_index, batch, None)
            except (BatchAssemblyError, TimeoutError) as e:
                # per-batch failure (its tasks already carry the exception): remember it and keep serving the other pools
                self.last_error = e
                continue
            yield item

    def _on_cuda(self) -> bool:
        return self.device is not None and torch.device(self.device).type == "cuda" and torch.cuda.is_available()

    def _prefetch_loop(self, out: "queue.Queue"):
        try:
            for item in self.iterate_minibatches_from_pools():
                while not self._stop.is_set():
                    try:
                        out.put(item, timeout=0.05)
                        break
                    except queue.Full:
                        continue
        except BaseException as e:  # noqa
            self.last_error = e
        finally:
            out.put(None)

    # ------------------------------------------------------------------ main loop
    def main(self):
        for pool in self.pools:
            pool.on_task = self._notify
            if not pool.is_alive():
                pool.start()
        if self.device is not None:
            for backend in self.expert_backends.values():
                backend.to(self.device)
        prefetched: "queue.Queue" = queue.Queue(maxsize=max(1, self.prefetch_batches))
        prefetcher = threading.Thread(target=self._prefetch_loop, args=(prefetched,), daemon=True)
        prefetcher.start()
        self.ready.set()
        with ThreadPoolExecutor(max_workers=max(1, self.sender_threads)) as senders:
            try:
                while True:
                    item = prefetched.get()
                    if item is None:
                        break
                    pool, batch_index, batch, ready = item
                    try:
                        if ready is not None:
                            torch.cuda.current_stream(self.device).wait_event(ready)
                        outputs = pool.process_func(*batch)
                        done = None
                        if self._on_cuda():
                            done = torch.cuda.Event()
                            done.record()
                        self.batches_processed += 1
                        self.samples_processed += len(outputs[0])
                        senders.submit(self.send_outputs_to_pool, pool, batch_index, outputs, done)
                    except BaseException as e:  # noqa: report to the clients instead of hanging them
                        self.last_error = e
                        pool.fail_batch(batch_index, e if isinstance(e, Exception) else RuntimeError(repr(e)))
                        if isinstance(e, KeyboardInterrupt):
                            break
            finally:
                self._stop.set()
                self._notify()
        prefetcher.join(timeout=1.0)
        for pool in self.pools:
            pool.join()

    def send_outputs_to_pool(self, pool: TaskPool, batch_index: int, outputs, done_event=None):
        try:
            if done_event is not None:
                done_event.synchronize()
            return pool.send_outputs_from_runtime(batch_index, outputs)
        except BaseException as e:  # noqa
            pool.fail_batch(batch_index, e if isinstance(e, Exception) else RuntimeError(repr(e)))

    def shutdown(self):
        self._stop.set()
        self._notify()
