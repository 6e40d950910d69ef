"""
Threading helpers, including ``run_and_await_k`` — the trainer-side straggler / fault-tolerance primitive
(parity: /root/reference/lib/utils/threading.py:7-125).
"""
import threading
import time
from concurrent.futures import Future, TimeoutError
from itertools import count
from typing import Callable, List, Optional, Sequence


def run_in_background(func: Callable, *args, **kwargs) -> Future:
    """Run func(*args, **kwargs) in a daemon thread; the returned Future carries its result or exception."""
    future = Future()

    def target():
        try:
            future.set_result(func(*args, **kwargs))
        except Exception as e:  # noqa: the exception is delivered through the future
            future.set_exception(e)

    threading.Thread(target=target, daemon=True).start()
    return future


def repeated(func: Callable, n_times: Optional[int] = None) -> Callable:
    """A callable that calls func forever (or n_times + 1 times, matching the reference's off-by-one)."""

    def repeat():
        for i in count():
            if n_times is not None and i > n_times:
                break
            func()

    return repeat


def add_event_callback(event: threading.Event, callback: Callable, timeout=None):
    """Call callback() from a helper thread once the event is set (or the timeout expires)."""
    threading.Thread(target=lambda: (event.wait(timeout), callback()), daemon=True).start()


class CountdownEvent(threading.Event):
    """An event that becomes set once it has been incremented :count_to: times."""

    def __init__(self, count_to: int, initial: int = 0):
        super().__init__()
        self.value, self.count_to = initial, count_to
        self.lock = threading.Lock()
        self.increment(by=0)

    def increment(self, by: int = 1) -> int:
        with self.lock:
            self.value += by
            if self.value >= self.count_to:
                super().set()
            else:
                super().clear()
            return self.value

    def clear(self):
        return self.increment(by=-self.value)


def await_first(*events: threading.Event, k: int = 1, timeout=None) -> bool:
    """Block until k of the events are set; afterwards all events are set (so helper threads terminate)."""
    done = CountdownEvent(count_to=k)
    for event in events:
        add_event_callback(event, callback=done.increment, timeout=timeout)
    if not done.wait(timeout=timeout):
        raise TimeoutError()
    for event in events:
        event.set()
    return True


def run_and_await_k(jobs: Sequence[Callable], k: int, timeout_after_k: float = 0, timeout_total: Optional[float] = None) -> List:
    """
    Start every job concurrently and wait until k of them succeeded (or so many failed that k can no longer succeed),
    then grant the stragglers :timeout_after_k: more seconds.

    :returns: one entry per job — its result, or the exception / TimeoutError it ended with
    :raises ValueError: more than len(jobs) - k jobs failed;  TimeoutError: timeout_total expired before k successes
    """
    jobs = list(jobs)
    assert 0 <= k <= len(jobs), "cannot await more jobs than were given"
    n = len(jobs)
    deadline = None if timeout_total is None else time.monotonic() + timeout_total
    cv = threading.Condition()
    state = dict(ok=0, failed=0)
    outcomes: List = [None] * n
    finished = [False] * n

    def worker(i, job):
        try:
            value, success = job(), True
        except Exception as e:  # noqa: failures are data here
            value, success = e, False
        with cv:
            outcomes[i], finished[i] = value, True
            state["ok" if success else "failed"] += 1
            cv.notify_all()

    for i, job in enumerate(jobs):
        threading.Thread(target=worker, args=(i, job), daemon=True).start()

    def decided():
        return state["ok"] >= k or state["failed"] > n - k

    with cv:
        while not decided():
            remaining = None if deadline is None else deadline - time.monotonic()
            if remaining is not None and remaining <= 0:
                break
            cv.wait(remaining)
        if state["ok"] >= k and not all(finished):
            # stragglers get a grace period
            grace_end = time.monotonic() + max(timeout_after_k, 0)
            if deadline is not None:
                grace_end = min(grace_end, deadline)
            while not all(finished):
                remaining = grace_end - time.monotonic()
                if remaining <= 0:
                    break
                cv.wait(remaining)
        results = [outcomes[i] if finished[i] else TimeoutError() for i in range(n)]
        enough, too_many_failed = state["ok"] >= k, state["failed"] > n - k

    if enough:
        return results
    if too_many_failed:
        raise ValueError("Could not get enough results: too many jobs failed.")
    raise TimeoutError("Could not get enough results: reached timeout_total.")
