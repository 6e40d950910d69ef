"""
TesseractServer — composition root of an expert server: DHT heartbeats + TCP front-end + device runtime
(API parity: /root/reference/lib/server/__init__.py:12-65).

The reference pre-forks ``conn_handler_processes`` acceptor PROCESSES and runs ~2 processes per expert; here the
front-end is ``conn_handler_processes`` acceptor THREADS in the server process (socket I/O and native framing release the
GIL), the pools are in-process, and one runtime thread owns the device.  ``start()`` blocks like the reference;
``run_in_background()`` / ``shutdown()`` exist for embedding and tests.  ``network=None`` disables discovery.
"""
import os
import threading
from socket import socket, AF_INET, SOCK_STREAM, SO_REUSEADDR, SOL_SOCKET, timeout
from typing import Dict, Optional

from ..runtime import TesseractRuntime, ExpertBackend
from .connection_handler import handle_connection
from .network_handler import NetworkHandlerThread


class TesseractServer:
    def __init__(self, network, expert_backends: Dict[str, ExpertBackend], addr="127.0.0.1", port: int = 8080,
                 conn_handler_processes: int = 1, update_period: int = 30, start=False, **kwargs):
        self.network, self.experts, self.update_period = network, expert_backends, update_period
        self.addr, self.port = addr, port
        self.conn_handlers = conn_handler_processes
        self.runtime = TesseractRuntime(self.experts, **kwargs)
        self._stop = threading.Event()
        self._threads = []
        self._network_thread: Optional[NetworkHandlerThread] = None
        self._sock: Optional[socket] = None
        self._background: Optional[threading.Thread] = None
        if start:
            self.start()

    # ------------------------------------------------------------------ lifecycle
    def start(self):
        """run the server in the calling thread; returns only after shutdown()"""
        self._threads = self.spawn_connection_handlers()  # binds the socket first: port=0 picks a free port
        if self.network:
            if not self.network.is_alive():
                self.network.start()
            self._network_thread = NetworkHandlerThread(experts=self.experts, network=self.network, addr=self.addr,
                                                        port=self.port, update_period=self.update_period)
            self._network_thread.start()
        try:
            self.runtime.main()
        finally:
            self._stop.set()
            if self._sock is not None:
                try:
                    self._sock.close()
                except OSError:
                    pass
            for thread in self._threads:
                thread.join(timeout=1.0)
            if self._network_thread is not None:
                self._network_thread.stop()

    def run_in_background(self, await_ready: bool = True, timeout: float = 30.0):
        self._background = threading.Thread(target=self.start, name=f"TesseractServer:{self.port}", daemon=True)
        self._background.start()
        if await_ready and not self.runtime.ready.wait(timeout):
            raise TimeoutError("server runtime did not become ready")
        return self

    def shutdown(self):
        self._stop.set()
        self.runtime.shutdown()
        if self._background is not None:
            self._background.join(timeout=5.0)

    @property
    def ready(self):
        return self.runtime.ready

    # ------------------------------------------------------------------ front-end
    def spawn_connection_handlers(self):
        sock = socket(AF_INET, SOCK_STREAM)
        sock.setsockopt(SOL_SOCKET, SO_REUSEADDR, 1)
        sock.bind(("", self.port))
        if self.port == 0:
            self.port = sock.getsockname()[1]
        sock.listen(1024)
        sock.settimeout(0.25)
        self._sock = sock
        threads = [threading.Thread(target=socket_loop, args=(sock, self.experts, self._stop), daemon=True,
                                    name=f"conn_handler_{i}") for i in range(max(1, self.conn_handlers))]
        for thread in threads:
            thread.start()
        return threads


def socket_loop(sock, experts, stop_event: Optional[threading.Event] = None):
    """accept connections, submit their task, reply; survives broken connections"""
    while stop_event is None or not stop_event.is_set():
        try:
            handle_connection(sock.accept(), experts)
        except KeyboardInterrupt:
            break
        except (timeout, BrokenPipeError, ConnectionResetError, NotImplementedError):
            continue
        except OSError:
            if stop_event is not None and stop_event.is_set():
                break
            continue
        except Exception as e:  # noqa: one bad connection must never cost the server an acceptor thread
            print(f"[TesseractServer] connection handler error ignored: {type(e).__name__}: {e}", flush=True)
            continue


__all__ = ["TesseractServer", "TesseractRuntime", "ExpertBackend", "socket_loop", "handle_connection",
           "NetworkHandlerThread"]
