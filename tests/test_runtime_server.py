"""TaskPool batching invariants, ExpertBackend semantics, TesseractServer + RemoteExpert over TCP (CPU)"""
import threading
import time

import pytest
import torch

import lah_b200 as lib
from lah_b200.models import FeedforwardBlock, TransformerEncoderLayer
from lah_b200.runtime import TaskPool


def make_pool(**kw):
    schema = (lib.BatchTensorProto(4),)
    return TaskPool(lambda x: (x * 2,), schema, schema, **kw)


def test_task_pool_batching_invariants():
    pool = make_pool(max_batch_size=8)
    xs = [torch.full((n, 4), float(i)) for i, n in enumerate([3, 3, 3, 2])]
    futures = [pool.submit_task(x) for x in xs]
    assert not pool.empty
    t_oldest = pool.priority
    index, (batch,) = pool.load_batch_to_runtime()
    # greedy: 3 + 3 < 8 -> takes the third task too (may overshoot max_batch_size by one task), order preserved
    assert batch.shape == (9, 4) and torch.equal(batch, torch.cat(xs[:3]))
    assert pool.priority > t_oldest  # priority now = timestamp of the 4th task
    pool.send_outputs_from_runtime(index, pool.process_func(batch))
    for fut, x in zip(futures[:3], xs[:3]):
        assert torch.equal(fut.result(timeout=1)[0], x * 2)
    index, (batch,) = pool.load_batch_to_runtime()
    assert batch.shape == (2, 4)
    pool.fail_batch(index, RuntimeError("device on fire"))
    with pytest.raises(RuntimeError):
        futures[3].result(timeout=1)
    assert pool.empty and pool.priority == float("inf")


def test_task_pool_min_batch_and_timeout():
    pool = make_pool(max_batch_size=16, min_batch_size=4, timeout=0.05)
    fut = pool.submit_task(torch.zeros(2, 4))
    with pytest.raises(TimeoutError):
        pool.form_batch()
    with pytest.raises(TimeoutError):
        fut.result(timeout=1)
    pool = make_pool(max_batch_size=16, min_batch_size=4, timeout=2.0)
    pool.submit_task(torch.zeros(2, 4))
    threading.Timer(0.05, lambda: pool.submit_task(torch.zeros(3, 4))).start()
    assert sum(TaskPool.get_task_size(t) for t in pool.form_batch()) == 5


def test_task_pool_blocks_when_full():
    pool = make_pool(max_batch_size=4, pool_size=1)
    pool.submit_task(torch.zeros(1, 4))
    done = threading.Event()
    threading.Thread(target=lambda: (pool.submit_task(torch.zeros(1, 4)), done.set()), daemon=True).start()
    assert not done.wait(0.1)
    pool.form_batch()
    assert done.wait(1.0)


def make_backend(uid="e", hid=16, **kw):
    expert = FeedforwardBlock(hid)
    return lib.ExpertBackend(name=uid, expert=expert, opt=torch.optim.Adam(expert.parameters(), lr=1e-2, amsgrad=True),
                             args_schema=(lib.BatchTensorProto(hid),), outputs_schema=lib.BatchTensorProto(hid),
                             max_batch_size=64, **kw)


def test_expert_backend_semantics():
    be = make_backend()
    assert list(be.state_dict().keys())[0] == "expert.layers.0.weight"  # reference checkpoint layout
    x = torch.randn(5, 16)
    (y1,) = be.forward(x)
    assert not y1.requires_grad and torch.allclose(y1, be.expert(x))
    before = be.expert.layers[0].weight.clone()
    (gx,) = be.backward(x, torch.ones(5, 16))
    assert gx.shape == x.shape and be.update_count == 1
    assert not torch.equal(before, be.expert.layers[0].weight)  # optimizer stepped right after backward
    (y2,) = be.forward(x)
    assert not torch.allclose(y1, y2)
    info = be.get_info()
    # the three keys of the reference + the negotiated raw-tensor-wire flag (ignored by reference clients)
    assert set(info) == {"forward_schema", "outputs_schema", "keyword_names", "tensor_wire"} and info["keyword_names"] == ()
    ckpt = be.checkpoint()
    be2 = make_backend()
    be2.load_checkpoint(ckpt)
    assert torch.equal(be2.forward(x)[0], y2)
    assert "max_exp_avg_sq" in ckpt["optimizer"]["state"][0]


def test_expert_backend_infers_output_schema_and_kwargs():
    class TwoInputs(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4, 6)

        def forward(self, x, scale):
            return self.lin(x) * scale, x.sum(-1, keepdim=True)

    m = TwoInputs()
    be = lib.ExpertBackend(name="two", expert=m, opt=torch.optim.SGD(m.parameters(), lr=0.1),
                           args_schema=(lib.BatchTensorProto(4),), kwargs_schema={"scale": lib.BatchTensorProto(1)},
                           max_batch_size=8)
    assert [p.size for p in be.outputs_schema] == [(None, 6), (None, 1)]
    grads = be.backward(torch.randn(3, 4), torch.ones(3, 1), torch.ones(3, 6), torch.ones(3, 1))
    assert [g.shape for g in grads] == [torch.Size([3, 4]), torch.Size([3, 1])]


def test_transformer_expert_is_trainable_and_does_not_mutate_input():
    block = TransformerEncoderLayer(32, nhead=4, dim_feedforward=64, dropout=0.0)
    be = lib.ExpertBackend(name="t", expert=block, opt=torch.optim.Adam(block.parameters()),
                           args_schema=(lib.BatchTensorProto(8, 32),), outputs_schema=lib.BatchTensorProto(8, 32),
                           max_batch_size=8)
    x = torch.randn(2, 8, 32)
    x0 = x.clone()
    (y,) = be.forward(x)
    assert y.shape == x.shape and torch.equal(x, x0)
    (gx,) = be.backward(x, torch.ones_like(y))  # the reference raises here (in-place transpose of a leaf)
    assert gx.shape == x.shape and be.update_count == 1


@pytest.fixture
def server():
    experts = {f"expert{i}": make_backend(f"expert{i}") for i in range(3)}
    srv = lib.TesseractServer(None, experts, port=0, conn_handler_processes=4, sender_threads=2)
    srv.run_in_background()
    yield srv
    srv.shutdown()


def test_remote_expert_forward_backward_over_tcp(server):
    remote = lib.RemoteExpert("expert1", "127.0.0.1", server.port)
    local = server.experts["expert1"].expert
    x = torch.randn(7, 16, requires_grad=True)
    expected = local(x.detach())
    out = remote(x)
    assert torch.allclose(out, expected, atol=1e-6)
    out.sum().backward()
    assert x.grad is not None and x.grad.shape == x.shape
    assert server.experts["expert1"].update_count == 1
    assert not torch.allclose(remote(x.detach()), expected)  # the server trained on our backward
    with pytest.raises(TypeError):
        remote(x, x)
    assert set(remote.info) == {"forward_schema", "outputs_schema", "keyword_names", "tensor_wire"}


def test_server_batches_concurrent_trainers_and_reports_errors(server):
    remote = lib.RemoteExpert("expert0", "127.0.0.1", server.port)
    remote.info
    xs = [torch.randn(4, 16) for _ in range(8)]
    outs = [None] * 8

    def call(i):
        outs[i] = remote(xs[i])

    threads = [threading.Thread(target=call, args=(i,)) for i in range(8)]
    [t.start() for t in threads]
    [t.join(10) for t in threads]
    local = server.experts["expert0"].expert
    for x, out in zip(xs, outs):
        assert torch.allclose(out, local(x), atol=1e-5)
    assert server.runtime.samples_processed >= 32
    # wrong shape: the reference hangs forever; we get the server-side error back
    bad = lib.RemoteExpert("expert0", "127.0.0.1", server.port)
    bad._info = remote.info
    with pytest.raises(lib.client.RemoteExpertError):
        bad(torch.randn(2, 5))
    with pytest.raises(lib.client.RemoteExpertError):
        lib.RemoteExpert("no_such_expert", "127.0.0.1", server.port).info


def test_runtime_serves_oldest_pool_first():
    b1, b2 = make_backend("a"), make_backend("b")
    rt = lib.TesseractRuntime({"a": b1, "b": b2})
    b2.forward_pool.submit_task(torch.zeros(1, 16))
    time.sleep(0.01)
    b1.forward_pool.submit_task(torch.zeros(1, 16))
    assert rt._next_pool(timeout=0) is b2.forward_pool  # older task wins (the reference would pick the newest)


def test_tensor_wire_roundtrip_and_protocol_negotiation(server):
    """raw-tensor frames (utils/tensor_wire.py): exact round trip; our client uses them against our server; a client that
    speaks the reference protocol ('fwd_' + torch.save) gets the same answer from the same server"""
    from lah_b200.utils import tensor_wire, Connection, PytorchSerializer
    tensors = (torch.randn(3, 5), torch.arange(6).view(2, 3), torch.randn(4).to(torch.bfloat16), torch.tensor([True, False]),
               torch.empty(0, 7), torch.randn(2, 2, 2).to(torch.float16))
    parts, total = tensor_wire.encode("some.uid", tensors)
    blob = b"".join(bytes(p) for p in parts)
    assert len(blob) == total
    uid, back = tensor_wire.decode(bytearray(blob))
    assert uid == "some.uid" and len(back) == len(tensors)
    for a, b in zip(tensors, back):
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    assert not tensor_wire.supported((torch.randn(2), "not a tensor"))
    # negotiated fast path == reference path
    remote = lib.RemoteExpert("expert2", "127.0.0.1", server.port)
    assert remote.info["tensor_wire"] == 1
    x = torch.randn(9, 16, requires_grad=True)
    y_fast = remote(x)
    with Connection.create("127.0.0.1", server.port) as c:      # what a reference client sends
        c.send_raw("fwd_", PytorchSerializer.dumps(("expert2", (x.detach(),))))
        header, message = c.recv_message()
    assert header == "rest"
    (y_ref,) = PytorchSerializer.loads(message)
    assert torch.allclose(y_fast, y_ref)
    y_fast.sum().backward()                                       # 'bwdT' works too (the server steps the expert)
    assert x.grad is not None and x.grad.shape == x.shape


def test_malformed_requests_do_not_kill_the_server():
    """a 'fwd_' with no tensors, mismatched shapes, garbage bytes, a truncated frame, a bad dtype code and an
    absurd length prefix each cost their author an error (or the connection) — the single acceptor thread and the runtime
    keep serving (the reference loses the expert's pool process / hangs its clients)."""
    import socket as pysocket
    import struct
    from lah_b200.utils import Connection, PytorchSerializer, tensor_wire
    experts = {"e": make_backend("e")}
    srv = lib.TesseractServer(None, experts, port=0, conn_handler_processes=1)   # ONE acceptor: losing it = dead server
    srv.run_in_background()
    try:
        def raw(payload: bytes):
            with pysocket.create_connection(("127.0.0.1", srv.port), timeout=5) as s:
                s.sendall(payload)
                s.settimeout(2)
                try:
                    return s.recv(1 << 16)
                except (pysocket.timeout, ConnectionResetError):
                    return b""

        def request(header, obj):
            with Connection.create("127.0.0.1", srv.port) as c:
                c.send_raw(header, PytorchSerializer.dumps(obj))
                return c.recv_message()

        h, _ = request("fwd_", ("e", ()))                                    # no tensors at all
        assert h == "err_"
        h, _ = request("fwd_", ("e", (torch.randn(2, 5),)))                  # wrong trailing shape
        assert h == "err_"
        h, _ = request("fwd_", ("e", (torch.randn(2, 16).double(),)))        # wrong dtype
        assert h == "err_"
        h, _ = request("bwd_", ("e", (torch.randn(2, 16), torch.randn(3, 16))))   # row counts disagree
        assert h == "err_"
        raw(b"fwd_" + (7).to_bytes(8, "big") + b"garbage")                   # bad pickle
        raw(b"fwdT" + (3).to_bytes(8, "big") + b"LAH")                       # truncated tensor frame
        frame = tensor_wire.MAGIC + struct.pack("<H", 1) + b"e" + struct.pack("<I", 1) + struct.pack("<BBqQ", 200, 1, 4, 16) + b"0" * 16
        raw(b"fwdT" + len(frame).to_bytes(8, "big") + frame)                 # dtype code out of range
        frame = tensor_wire.MAGIC + struct.pack("<H", 1) + b"e" + struct.pack("<I", 1) + struct.pack("<BBqQ", 0, 1, 1 << 40, 16) + b"0" * 16
        raw(b"fwdT" + len(frame).to_bytes(8, "big") + frame)                 # dims claim 4 TiB, 16 bytes sent
        raw(b"fwd_" + (1 << 62).to_bytes(8, "big"))                          # absurd length prefix
        # still alive and correct
        remote = lib.RemoteExpert("e", "127.0.0.1", srv.port)
        x = torch.randn(3, 16)
        assert torch.allclose(remote(x), experts["e"].expert(x), atol=1e-6)
        assert all(t.is_alive() for t in srv._threads)
    finally:
        srv.shutdown()
    with pytest.raises(ValueError):
        tensor_wire.decode(bytearray(b"LAHT\x01\x00e\xff\xff\xff\xff"))      # 4e9 tensors announced
