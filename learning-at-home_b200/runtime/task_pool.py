"""
TaskPool — groups requests from many trainers into batches for one expert (dynamic batching across trainers).

API parity with /root/reference/lib/runtime/task_pool.py:22-237 (constructor signature, ``submit_task``,
``form_batch``, ``load_batch_to_runtime``, ``send_outputs_from_runtime``, ``priority``, ``empty``), different design:

* The reference runs every pool as an OS process (2 per expert, ~131 processes per GPU) and ships batches through
  POSIX shared memory and pipes.  Here a pool is a thread-safe in-process queue: connection-handler threads call
  ``submit_task``, the runtime thread calls ``load_batch_to_runtime`` which forms the batch, assembles it into a
  reusable PINNED staging buffer (native multi-threaded row gather, csrc/host_runtime.cpp) and issues ONE async H2D copy.
* ``priority`` is the submission time of the OLDEST waiting task; the runtime serves the pool with the smallest value
  (oldest-waiting-first).  The reference intends this but actually picks the newest (SURVEY.md §2.3 "quirk").
* batch-size semantics are kept: greedy batches, at least ``min_batch_size`` rows, may overshoot ``max_batch_size`` by
  the last task; with ``timeout`` set, an incomplete batch fails its tasks with TimeoutError.
"""
import threading
import time
import uuid
from collections import deque, namedtuple
from concurrent.futures import Future
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from ..utils import BatchTensorProto

Task = namedtuple("Task", ("future", "args", "timestamp"))


class BatchAssemblyError(RuntimeError):
    """a formed batch could not be assembled; its tasks have already been failed (per-batch error, not fatal)"""


class TaskPoolBase:
    """Common interface between pools and TesseractRuntime (the runtime only needs these members)."""

    def __init__(self, process_func: callable):
        self.process_func = process_func
        self._priority = float("inf")

    def submit_task(self, *args: torch.Tensor) -> Future:
        raise NotImplementedError()

    def form_batch(self, *args, **kwargs) -> List[Task]:
        raise NotImplementedError()

    def iterate_minibatches(self, *args, **kwargs):
        while True:
            yield self.form_batch(*args, **kwargs)

    @property
    def priority(self) -> float:
        """submission time of the oldest waiting task (smaller = more urgent); +inf when the pool is empty"""
        return self._priority

    @priority.setter
    def priority(self, value):
        self._priority = float(value)

    @property
    def empty(self) -> bool:
        raise NotImplementedError()


class TaskPool(TaskPoolBase):
    def __init__(self, process_func: callable, inputs_schema: Tuple[BatchTensorProto, ...],
                 outputs_schema: Tuple[BatchTensorProto, ...], max_batch_size: int, min_batch_size: int = 1,
                 timeout: Optional[float] = None, pool_size: Optional[int] = None, prefetch_batches: int = 1,
                 uid=None, shm_manager=None, array_headers=None, start: bool = False):
        """
        :param process_func: called by the runtime on every formed batch: process_func(*tensors) -> sequence of tensors
        :param max_batch_size: stop adding tasks once a batch has this many rows (may overshoot by one task)
        :param min_batch_size: wait for at least this many rows
        :param timeout: seconds to wait for the next task while the batch is still below min_batch_size
        :param pool_size: maximum number of waiting tasks; submit_task blocks when the pool is full
        :param shm_manager, array_headers: accepted for source compatibility with the reference; unused (no shm hop)
        """
        super().__init__(process_func)
        self.inputs_schema, self.outputs_schema = list(inputs_schema), list(outputs_schema)
        self.max_batch_size, self.min_batch_size, self.timeout = max_batch_size, min_batch_size, timeout
        self.pool_size, self.prefetch_batches = pool_size or 0, prefetch_batches
        self.uid = uid or uuid.uuid4()
        self._tasks = deque()
        self._lock = threading.Lock()
        self._not_empty = threading.Condition(self._lock)
        self._not_full = threading.Condition(self._lock)
        self._pending: Dict[int, List[Task]] = {}
        self._next_batch_index = 0
        self._staging: Dict[Tuple[int, int], torch.Tensor] = {}
        self._staging_free = None  # CUDA event: H2D copies out of the staging buffers have completed
        self._alive = False
        self.on_task = None  # set by the runtime: wakes its scheduler when a task arrives
        if start:
            self.start()

    # ------------------------------------------------------------------ process-like lifecycle (API parity)
    def start(self):
        self._alive = True

    def is_alive(self) -> bool:
        return self._alive

    def join(self, timeout=None):
        self._alive = False

    # ------------------------------------------------------------------ producer side (connection handlers)
    def validate_task(self, args) -> None:
        """reject a request that cannot be batched with the others BEFORE it enters the queue (a malformed request must
        cost its author an 'err_' reply, not the server its runtime): argument count, tensor-ness, dtype, trailing shape
        and equal row counts are checked against ``inputs_schema``"""
        if len(args) != len(self.inputs_schema):
            raise ValueError(f"expected {len(self.inputs_schema)} input tensors, got {len(args)}")
        rows = None
        for i, (arg, proto) in enumerate(zip(args, self.inputs_schema)):
            if not isinstance(arg, torch.Tensor):
                raise TypeError(f"input {i} is not a tensor")
            want = tuple(proto.size[1:]) if getattr(proto, "size", None) is not None else None
            if arg.dim() < 1 or (want is not None and tuple(arg.shape[1:]) != want):
                raise ValueError(f"input {i} has shape {tuple(arg.shape)}, expected [batch, {', '.join(map(str, want or ()))}]")
            if getattr(proto, "dtype", None) is not None and arg.dtype != proto.dtype:
                raise TypeError(f"input {i} has dtype {arg.dtype}, expected {proto.dtype}")
            if rows is None:
                rows = arg.shape[0]
            elif arg.shape[0] != rows:
                raise ValueError("all inputs of a request must have the same number of rows")
        if not rows:
            raise ValueError("empty request")

    def submit_task(self, *args: torch.Tensor) -> Future:
        future = Future()
        try:
            self.validate_task(args)
        except Exception as e:  # delivered to the author through the future (-> 'err_' reply)
            future.set_exception(e)
            return future
        task = Task(future, args, time.time())
        with self._lock:
            while self.pool_size and len(self._tasks) >= self.pool_size:
                self._not_full.wait()
            self._tasks.append(task)
            self._priority = self._tasks[0].timestamp
            self._not_empty.notify()
        if self.on_task is not None:
            self.on_task(self)
        return future

    # ------------------------------------------------------------------ batching
    @staticmethod
    def get_task_size(task: Task) -> int:
        """rows contributed by a task (the batching unit)"""
        return len(task.args[0]) if task.args else 1

    @property
    def empty(self) -> bool:
        return not self._tasks

    def form_batch(self, block: bool = True) -> List[Task]:
        batch, total = [], 0
        with self._lock:
            while total < self.max_batch_size:
                if not self._tasks:
                    if total >= self.min_batch_size or (not block and not batch):
                        break
                    if not self._not_empty.wait(self.timeout) and not self._tasks:
                        exc = TimeoutError(f"Timeout reached but batch doesn't contain >={self.min_batch_size} elements yet.")
                        for task in batch:
                            task.future.set_exception(exc)
                        raise exc
                    continue
                task = self._tasks.popleft()
                if task.future.set_running_or_notify_cancel():
                    batch.append(task)
                    total += self.get_task_size(task)
            self._priority = self._tasks[0].timestamp if self._tasks else float("inf")
            self._not_full.notify_all()
        return batch

    # ------------------------------------------------------------------ runtime side
    def _staging_buffer(self, index: int, rows: int, proto: BatchTensorProto, like: torch.Tensor, pin: bool):
        """reusable (pinned) host buffer with capacity rounded up to a power of two, one per input slot"""
        capacity = 1 << max(rows - 1, 0).bit_length()
        key = (index, capacity)
        buf = self._staging.get(key)
        if buf is None:
            buf = torch.empty((capacity, *like.shape[1:]), dtype=like.dtype, pin_memory=pin)
            self._staging = {k: v for k, v in self._staging.items() if k[0] != index}  # keep one buffer per slot
            self._staging[key] = buf
        return buf[:rows]

    def load_batch_to_runtime(self, timeout=None, device=None) -> Tuple[Any, List[torch.Tensor]]:
        """form the next batch, assemble it and start moving it to :device:; returns (batch_index, tensors)"""
        if timeout is not None:
            deadline = time.time() + timeout
            while self.empty:
                if time.time() > deadline:
                    raise TimeoutError()
                time.sleep(0.0005)
        tasks = self.form_batch()
        batch_index = self._next_batch_index
        self._next_batch_index += 1
        self._pending[batch_index] = tasks
        try:
            return batch_index, self._assemble(tasks, device)
        except Exception as e:   # a batch that cannot be assembled fails ITS tasks only; the runtime keeps serving
            self.fail_batch(batch_index, e)
            raise BatchAssemblyError(str(e)) from e

    def _assemble(self, tasks, device):
        rows = sum(map(self.get_task_size, tasks))
        to_cuda = device is not None and torch.device(device).type == "cuda"
        batch = []
        from ..ops import host
        if to_cuda and self._staging_free is not None:
            self._staging_free.synchronize()  # the previous batch's H2D copies have left the pinned staging buffers
        for i, proto in enumerate(self.inputs_schema):
            parts = [task.args[i] for task in tasks]
            if to_cuda:
                # assemble into a reusable PINNED buffer (native threaded gather), then one async H2D copy
                staged = self._staging_buffer(i, rows, proto, parts[0], pin=True)
                host.gather_rows(parts, staged)
                tensor = staged.to(device, non_blocking=True)
            else:
                # host execution: the batch must own its memory (the runtime may still hold the previous batch)
                tensor = parts[0] if len(parts) == 1 else host.gather_rows(
                    parts, torch.empty((rows, *parts[0].shape[1:]), dtype=parts[0].dtype))
                if device is not None:
                    tensor = tensor.to(device)
            batch.append(tensor)
        if to_cuda:
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream(device))
        return batch

    def send_outputs_from_runtime(self, batch_index: int, batch_outputs: Sequence):
        """split the outputs of a processed batch by task and resolve the tasks' futures"""
        tasks = self._pending.pop(batch_index)
        sizes = [self.get_task_size(task) for task in tasks]
        outputs = [out if isinstance(out, torch.Tensor) else torch.as_tensor(out) for out in batch_outputs]
        outputs = [out.detach().to("cpu") if out.is_cuda else out.detach() for out in outputs]
        offset = 0
        for task, size in zip(tasks, sizes):
            task.future.set_result(tuple(out[offset: offset + size].clone() for out in outputs))
            offset += size

    def fail_batch(self, batch_index: int, exception: BaseException):
        """propagate a runtime error to the authors of the batch (the reference leaves its clients hanging)"""
        for task in self._pending.pop(batch_index, []):
            task.future.set_exception(exception)
