"""
TCP message framing of the out-of-box (CPU / TCP) fallback path.

Wire format (parity: /root/reference/lib/utils/connection.py:6-54):
    message := header(4 ASCII bytes) length(8 bytes, big-endian) payload(length bytes)

Differences from the reference implementation (same bytes on the wire):
  * receives go straight into one pre-sized bytearray with ``recv_into`` (the reference loops over 2 KiB ``recv`` calls
    and joins the chunks: 4096 Python iterations for an 8 MiB tensor),
  * sends use ``sendall`` on a single gathered buffer, TCP_NODELAY is set,
  * short reads of the header / length fields are handled (the reference assumes they arrive whole).
"""
import socket
from contextlib import AbstractContextManager
from typing import Tuple

HEADER_SIZE = 4
LENGTH_SIZE = 8


class Connection(AbstractContextManager):
    header_size = HEADER_SIZE  # number of characters in all headers
    payload_length_size = LENGTH_SIZE  # number of bytes used to encode payload length

    __slots__ = ("conn", "addr")

    def __init__(self, conn: socket.socket, addr: Tuple[str, int]):
        self.conn, self.addr = conn, addr

    @staticmethod
    def create(host: str, port: int, timeout=None) -> "Connection":
        sock = socket.create_connection((host, port), timeout=timeout)
        sock.settimeout(None)
        try:
            sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        except OSError:
            pass
        return Connection(sock, (host, port))

    # ---------------------------------------------------------------- send
    def send_raw(self, header: str, content: bytes) -> None:
        head = header.encode()
        assert len(head) == self.header_size, f"header must be {self.header_size} ASCII characters"
        prefix = head + len(content).to_bytes(self.payload_length_size, byteorder="big")
        if len(content) <= 65536:
            self.conn.sendall(prefix + content)
        else:
            self.conn.sendall(prefix)
            self.conn.sendall(content)

    def send_parts(self, header: str, parts, total_length: int) -> None:
        """one message whose payload is the concatenation of ``parts`` (bytes / memoryviews), sent without gathering them"""
        head = header.encode()
        assert len(head) == self.header_size, f"header must be {self.header_size} ASCII characters"
        self.conn.sendall(head + total_length.to_bytes(self.payload_length_size, byteorder="big"))
        for part in parts:
            self.conn.sendall(part)

    # ---------------------------------------------------------------- receive
    #: payloads up to this size are received into one pre-sized buffer; larger ones grow the buffer (doubling) as the bytes
    #: actually ARRIVE, so a peer that merely announces gigabytes cannot make this process allocate them
    eager_alloc = 64 << 20

    def _recv_exact(self, nbytes: int) -> bytearray:
        buf = bytearray(min(nbytes, self.eager_alloc))
        got = 0
        while got < nbytes:
            if got == len(buf):                                   # large message: the announced bytes keep coming -> grow
                buf.extend(bytes(min(nbytes - got, len(buf))))
            with memoryview(buf) as view:                         # (released before the next extend: no exported buffer)
                n = self.conn.recv_into(view[got:], len(buf) - got)
            if n == 0:
                raise RuntimeError("socket connection broken")
            got += n
        return buf

    def recv_header(self) -> str:
        return self._recv_exact(self.header_size).decode()

    #: largest payload a peer may announce (the length prefix is attacker-controlled: never allocate it blindly)
    max_payload = 1 << 32

    def _recv_length(self) -> int:
        length = int.from_bytes(self._recv_exact(self.payload_length_size), byteorder="big")
        if length > self.max_payload:
            raise ValueError(f"announced payload of {length} bytes exceeds the limit of {self.max_payload}")
        return length

    def recv_raw(self, max_package: int = 0) -> bytes:
        """:param max_package: kept for API compatibility with the reference; ignored"""
        return bytes(self._recv_exact(self._recv_length()))

    def recv_buffer(self) -> bytearray:
        """payload as the receive buffer itself (no copy): for zero-copy decoders (utils/tensor_wire.py)"""
        return self._recv_exact(self._recv_length())

    def recv_message(self) -> Tuple[str, bytes]:
        return self.recv_header(), self.recv_raw()

    def close(self):
        try:
            self.conn.close()
        except OSError:
            pass

    def __exit__(self, *exc_info):
        self.close()
