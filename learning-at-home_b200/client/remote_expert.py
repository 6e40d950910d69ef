"""
RemoteExpert — an ``nn.Module`` whose forward/backward run on a TesseractServer (TCP/CPU fallback path; the in-box fast
path never goes through here).  API + wire protocol parity: /root/reference/lib/client/remote_expert.py:9-76 and
SURVEY.md appendix B ('info' / 'fwd_' / 'bwd_' requests, 'rest' replies, torch.save payloads).

Additions: a server-side exception is transported back and re-raised (reply header 'err_'; the reference leaves the
client hanging), and an optional per-call ``timeout``.
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from ..utils import Connection, DUMMY, PytorchSerializer, nested_compare, nested_flatten, nested_pack
from ..utils import tensor_wire


class RemoteExpertError(RuntimeError):
    pass


def _call(host: str, port: int, header: str, payload, timeout=None):
    with Connection.create(host, port, timeout=timeout) as connection:
        if timeout is not None:
            connection.conn.settimeout(timeout)
        connection.send_raw(header, PytorchSerializer.dumps(payload))
        reply_header, message = connection.recv_message()
    result = PytorchSerializer.loads(message)
    if reply_header == "err_":
        raise RemoteExpertError(f"{header.strip('_')} on {host}:{port} failed on the server: {result}")
    return result


def _call_tensors(host: str, port: int, header: str, uid: str, tensors, timeout=None, fast: bool = False):
    """'fwd_' / 'bwd_' request; with ``fast`` the negotiated raw-tensor frames are used (no torch.save on either side)"""
    if not (fast and tensor_wire.supported(tensors)):
        return _call(host, port, header, (uid, tuple(tensors)), timeout)
    fast_header = {"fwd_": "fwdT", "bwd_": "bwdT"}[header]
    with Connection.create(host, port, timeout=timeout) as connection:
        if timeout is not None:
            connection.conn.settimeout(timeout)
        connection.send_parts(fast_header, *tensor_wire.encode(uid, tensors))
        reply_header = connection.recv_header()
        if reply_header == tensor_wire.REPLY_HEADER:
            return tensor_wire.decode(connection.recv_buffer())[1]
        result = PytorchSerializer.loads(connection.recv_raw())
    if reply_header == "err_":
        raise RemoteExpertError(f"{header.strip('_')} on {host}:{port} failed on the server: {result}")
    return result


class RemoteExpert(nn.Module):
    """
    :param uid: unique expert identifier
    :param host: hostname where the TesseractServer listens
    :param port: its port
    """

    def __init__(self, uid, host="127.0.0.1", port=8080, timeout: Optional[float] = None):
        super().__init__()
        self.uid, self.host, self.port, self.timeout = uid, host, port, timeout
        self._info = None

    def forward(self, *args, **kwargs):
        info = self.info
        assert len(kwargs) == len(info["keyword_names"]), f"Keyword args should be {info['keyword_names']}"
        kwargs = {key: kwargs[key] for key in info["keyword_names"]}  # server-side order
        forward_inputs = (args, kwargs)
        if not nested_compare(forward_inputs, info["forward_schema"]):
            raise TypeError("Inputs do not match expert input schema. Did you pass the right number of parameters?")
        fast = bool(info.get("tensor_wire"))   # only servers of this package advertise the raw-tensor extension
        flat_outputs = _RemoteModuleCall.apply(DUMMY, self.uid, self.host, self.port, (self.timeout, fast),
                                               *nested_flatten(forward_inputs))
        return nested_pack(flat_outputs, structure=info["outputs_schema"])

    @property
    def info(self):
        if self._info is None:
            self._info = _call(self.host, self.port, "info", self.uid, self.timeout)
        return self._info

    def extra_repr(self):
        return f"uid={self.uid}, host={self.host}, port={self.port}"

    # experts are used as dict keys / set members by GatingFunction
    def __hash__(self):
        return hash((self.uid, self.host, self.port))

    def __eq__(self, other):
        return isinstance(other, RemoteExpert) and (self.uid, self.host, self.port) == (other.uid, other.host, other.port)


class _RemoteModuleCall(torch.autograd.Function):
    """autograd bridge: forward = 'fwd_' RPC, backward = 'bwd_' RPC carrying the saved inputs + output gradients
    (the server is stateless between the two calls)"""

    @staticmethod
    def forward(ctx, dummy, uid, host, port, timeout, *inputs):
        inputs = tuple(t.detach() for t in inputs)
        timeout, fast = timeout if isinstance(timeout, tuple) else (timeout, False)
        ctx.uid, ctx.host, ctx.port, ctx.timeout, ctx.fast = uid, host, port, timeout, fast
        ctx.save_for_backward(*inputs)
        outputs = _call_tensors(host, port, "fwd_", uid, tuple(t.cpu() for t in inputs), timeout, fast)
        device = inputs[0].device if inputs else torch.device("cpu")
        return tuple(t.to(device) for t in outputs)

    @staticmethod
    def backward(ctx, *grad_outputs) -> Tuple[Optional[torch.Tensor], ...]:
        inputs = ctx.saved_tensors
        payload = tuple(t.detach().cpu() for t in nested_flatten((inputs, grad_outputs)))
        grad_inputs = _call_tensors(ctx.host, ctx.port, "bwd_", ctx.uid, payload, ctx.timeout, ctx.fast)
        device = inputs[0].device if inputs else torch.device("cpu")
        return (DUMMY, None, None, None, None, *(g.to(device) for g in grad_inputs))
