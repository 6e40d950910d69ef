"""Heartbeat thread: re-declares the server's experts every update_period seconds
(parity: /root/reference/lib/server/network_handler.py:7-20)."""
import threading


class NetworkHandlerThread(threading.Thread):
    def __init__(self, experts, network, update_period: int = 5, addr: str = "127.0.0.1", port: int = 8080):
        super().__init__(daemon=True)
        self.experts, self.network = experts, network
        self.update_period, self.addr, self.port = update_period, addr, port
        self._stop_event = threading.Event()

    def run(self) -> None:
        while not self._stop_event.is_set():
            try:
                self.network.declare_experts(list(self.experts.keys()), self.addr, self.port)
            except Exception:  # noqa: a failed heartbeat must not kill the server; it is retried next period
                pass
            self._stop_event.wait(self.update_period)

    def stop(self):
        self._stop_event.set()
