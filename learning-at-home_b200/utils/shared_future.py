"""
A Future whose two ends live in different processes, linked by a multiprocessing Pipe
(parity: /root/reference/lib/utils/shared_future.py:7-105).
"""
import multiprocessing as mp
import multiprocessing.connection
from concurrent.futures import CancelledError, Future


class SharedFuture(Future):
    STATES = ("pending", "running", "cancelled", "finished", "exception")
    STATE_PENDING, STATE_RUNNING, STATE_CANCELLED, STATE_FINISHED, STATE_EXCEPTION = STATES
    _TERMINAL = (STATE_CANCELLED, STATE_FINISHED, STATE_EXCEPTION)

    def __init__(self, connection: mp.connection.Connection):
        """Prefer SharedFuture.make_pair(); :connection: is this end of the pipe."""
        # NOTE: deliberately NOT calling Future.__init__: no condition variable, so the object pickles across processes
        self.connection = connection
        self.state = self.STATE_PENDING
        self._result = None
        self._exception = None

    @classmethod
    def make_pair(cls):
        """Two linked futures: whatever one end sets, the other end observes."""
        end_a, end_b = mp.Pipe()
        return cls(end_a), cls(end_b)

    # ------------------------------------------------------------------ receiving side
    def _sync(self, timeout):
        """Pull state updates from the pipe until a terminal state arrives (or time out)."""
        while self.state not in self._TERMINAL:
            if not self.connection.poll(timeout):
                raise TimeoutError()
            try:
                state, payload = self.connection.recv()
            except (BrokenPipeError, EOFError, ConnectionResetError) as e:
                state, payload = self.STATE_EXCEPTION, e
            if state not in self.STATES or state == self.STATE_PENDING:
                raise ValueError(f"unexpected future state on the wire: {state!r}")
            self.state = state
            if state == self.STATE_FINISHED:
                self._result = payload
            elif state == self.STATE_EXCEPTION:
                self._exception = payload
            if state == self.STATE_RUNNING:
                continue

    def result(self, timeout=None):
        self._sync(timeout)
        if self.state == self.STATE_FINISHED:
            return self._result
        if self.state == self.STATE_EXCEPTION:
            raise self._exception
        raise CancelledError()

    def exception(self, timeout=None):
        self._sync(timeout)
        return self._exception

    # ------------------------------------------------------------------ sending side
    def _publish(self, state, payload) -> bool:
        try:
            self.connection.send((state, payload))
            return True
        except (BrokenPipeError, OSError):
            return False

    def set_result(self, result) -> bool:
        self.state, self._result = self.STATE_FINISHED, result
        return self._publish(self.STATE_FINISHED, result)

    def set_exception(self, exception: BaseException) -> bool:
        self.state, self._exception = self.STATE_EXCEPTION, exception
        return self._publish(self.STATE_EXCEPTION, exception)

    def set_running_or_notify_cancel(self) -> bool:
        return True

    # ------------------------------------------------------------------ misc Future API
    def done(self) -> bool:
        self._poll_quietly()
        return self.state in self._TERMINAL

    def running(self) -> bool:
        self._poll_quietly()
        return self.state == self.STATE_RUNNING

    def cancelled(self) -> bool:
        return self.state == self.STATE_CANCELLED

    def cancel(self):
        raise NotImplementedError("SharedFuture cannot be cancelled")

    def add_done_callback(self, callback):
        raise NotImplementedError("SharedFuture does not support callbacks")

    def _poll_quietly(self):
        try:
            self._sync(timeout=0)
        except (TimeoutError, OSError, ValueError):
            pass

    def __repr__(self):
        self._poll_quietly()
        if self.state == self.STATE_FINISHED:
            return f"<SharedFuture at 0x{id(self):x} state=finished returned {type(self._result).__name__}>"
        if self.state == self.STATE_EXCEPTION:
            return f"<SharedFuture at 0x{id(self):x} state=finished raised {type(self._exception).__name__}>"
        return f"<SharedFuture at 0x{id(self):x} state={self.state}>"
