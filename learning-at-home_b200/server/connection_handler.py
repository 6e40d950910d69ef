"""
One request per TCP connection: 'fwd_' / 'bwd_' / 'info' -> reply 'rest' (parity:
/root/reference/lib/server/connection_handler.py:8-29).  Unlike the reference, a failure while serving the request is
reported to the client with an 'err_' reply instead of silence.
"""
from socket import socket
from typing import Dict, Tuple

from ..runtime.expert_backend import ExpertBackend
from ..utils import Connection, PytorchSerializer


def handle_connection(connection_tuple: Tuple[socket, str], experts: Dict[str, ExpertBackend], task_timeout=None):
    with Connection(*connection_tuple) as connection:
        try:
            header = connection.recv_header()
            payload = PytorchSerializer.loads(connection.recv_raw())
        except (RuntimeError, OSError, EOFError):
            return  # client went away
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
            connection.send_raw(reply_header, PytorchSerializer.dumps(response))
        except (RuntimeError, OSError):
            pass
