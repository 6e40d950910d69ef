"""
One request per TCP connection: 'fwd_' / 'bwd_' / 'info' -> reply 'rest' (parity:
/root/reference/lib/server/connection_handler.py:8-29).  Unlike the reference, a failure while serving the request is
reported to the client with an 'err_' reply instead of silence.
"""
from socket import socket
from typing import Dict, Tuple

from ..runtime.expert_backend import ExpertBackend
from ..utils import Connection, PytorchSerializer
from ..utils import tensor_wire


def handle_connection(connection_tuple: Tuple[socket, str], experts: Dict[str, ExpertBackend], task_timeout=None):
    with Connection(*connection_tuple) as connection:
        try:
            header = connection.recv_header()
            fast = header in tensor_wire.REQUEST_HEADERS   # raw-tensor frames (negotiated extension, utils/tensor_wire.py)
            if fast:
                payload = tensor_wire.decode(connection.recv_buffer())
                header = tensor_wire.REQUEST_HEADERS[header]
            else:
                payload = PytorchSerializer.loads(connection.recv_raw())
        except Exception:  # noqa: client went away / garbage bytes (bad pickle, truncated frame, oversized length prefix, ...)
            return      # drop the connection; the acceptor thread lives on
        try:
            if header == "fwd_":
                uid, inputs = payload
                response = experts[uid].forward_pool.submit_task(*inputs).result(task_timeout)
            elif header == "bwd_":
                uid, inputs_and_grad_outputs = payload
                response = experts[uid].backward_pool.submit_task(*inputs_and_grad_outputs).result(task_timeout)
            elif header == "info":
                response = experts[payload].get_info()
            else:
                raise NotImplementedError(f"Unknown header: {header}")
            reply_header = "rest"
        except BaseException as e:  # noqa: delivered to the client
            reply_header, response = "err_", f"{type(e).__name__}: {e}"
        try:
            if fast and reply_header == "rest" and tensor_wire.supported(response):
                connection.send_parts(tensor_wire.REPLY_HEADER, *tensor_wire.encode("", response))
            else:
                connection.send_raw(reply_header, PytorchSerializer.dumps(response))
        except (RuntimeError, OSError):
            pass
