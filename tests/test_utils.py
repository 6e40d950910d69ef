"""CPU unit tests of lah_b200.utils (reference: /root/reference/lib/utils/*)"""
import collections
import multiprocessing as mp
import socket
import threading
import time

import numpy as np
import pytest
import torch

import lah_b200 as lib
from lah_b200 import utils


def test_nested_roundtrip_and_order():
    Point = collections.namedtuple("Point", ["x", "y"])
    struct = (1, [2, 3], {"b": 4, "a": 5}, Point(6, 7))
    flat = list(utils.nested_flatten(struct))
    assert flat == [1, 2, 3, 5, 4, 6, 7]  # dict values in SORTED key order
    packed = utils.nested_pack(flat, struct)
    assert packed == (1, [2, 3], {"a": 5, "b": 4}, Point(6, 7)) and isinstance(packed[3], Point)
    assert utils.nested_compare(struct, packed)
    assert not utils.nested_compare((1, 2), (1, 2, 3)) and not utils.nested_compare({"a": 1}, {"b": 1})
    assert not utils.nested_compare([1], (1,))
    assert utils.nested_map(lambda a, b: a + b, struct, struct) == (2, [4, 6], {"a": 10, "b": 8}, Point(12, 14))
    with pytest.raises(ValueError):
        utils.nested_map(lambda a, b: a + b, (1, 2), (1, 2, 3))


def test_protos():
    proto = utils.BatchTensorProto(5, 7, dtype=torch.float16)
    assert proto.size == (None, 5, 7) and proto.shape == (None, 5, 7)
    assert proto.make_empty(3).shape == (3, 5, 7) and proto.make_empty(3).dtype == torch.float16
    assert utils.BatchTensorProto((5, 7)).size == (None, 5, 7)
    t = torch.zeros(4, 9, dtype=torch.int64)
    assert utils.BatchTensorProto.from_tensor(t).size == (None, 9)
    tp = utils.TensorProto.from_tensor(t)
    assert tp.size == (4, 9) and tp.make_empty().shape == (4, 9)
    arr = np.arange(12, dtype=np.float32).reshape(3, 4)
    ap = utils.ArrayProto.from_array(arr)
    assert ap.nbytes == 48 and ap.make_empty().shape == (3, 4)
    buf = bytearray(48)
    view = ap.make_from_buffer(buf)
    view[...] = arr
    assert np.array_equal(np.frombuffer(buf, dtype=np.float32).reshape(3, 4), arr)
    assert torch.equal(tp.convert_array_to_tensor(np.ones((4, 9), dtype=np.int64)), torch.ones(4, 9, dtype=torch.int64))
    assert utils.DUMMY_BATCH_SIZE == 3 and utils.DUMMY.requires_grad and utils.DUMMY.numel() == 0


def test_serializers():
    obj = ("uid", (torch.arange(6).view(2, 3), utils.BatchTensorProto(3)))
    back = utils.PytorchSerializer.loads(utils.PytorchSerializer.dumps(obj))
    assert back[0] == "uid" and torch.equal(back[1][0], obj[1][0]) and back[1][1] == obj[1][1]
    assert utils.PickleSerializer.loads(utils.PickleSerializer.dumps({"a": 1})) == {"a": 1}
    assert utils.JoblibSerializer.loads(utils.JoblibSerializer.dumps([1, 2])) == [1, 2]


def test_connection_framing_roundtrip():
    a, b = socket.socketpair()
    ca, cb = utils.Connection(a, ("local", 0)), utils.Connection(b, ("local", 0))
    payload = bytes(np.random.randint(0, 255, size=3_000_000, dtype=np.uint8))
    t = threading.Thread(target=lambda: ca.send_raw("fwd_", payload))
    t.start()
    header, got = cb.recv_message()
    t.join()
    assert header == "fwd_" and got == payload
    # exact bytes on the wire: 4-char header, 8-byte big-endian length (reference framing)
    ca.send_raw("info", b"xyz")
    raw = b.recv(64)
    assert raw == b"info" + (3).to_bytes(8, "big") + b"xyz"
    ca.close()
    with pytest.raises(RuntimeError):
        cb.recv_message()
    cb.close()


def _fulfil(fut, delay, value=None, exc=None):
    time.sleep(delay)
    fut.set_exception(exc) if exc is not None else fut.set_result(value)


def test_shared_future_result_exception_timeout():
    f1, f2 = utils.SharedFuture.make_pair()
    with pytest.raises(TimeoutError):
        f2.result(timeout=0.05)
    threading.Thread(target=_fulfil, args=(f1, 0.05, {"x": 1})).start()
    assert f2.result(timeout=2) == {"x": 1} and f2.done()
    g1, g2 = utils.SharedFuture.make_pair()
    g1.set_exception(ValueError("boom"))
    with pytest.raises(ValueError):
        g2.result(timeout=1)
    assert isinstance(g2.exception(), ValueError)
    # across processes
    h1, h2 = utils.SharedFuture.make_pair()
    p = mp.get_context("fork").Process(target=_fulfil, args=(h1, 0.0, 42))
    p.start()
    assert h2.result(timeout=5) == 42
    p.join()


def test_run_and_await_k():
    def ok(v, delay=0.0):
        def job():
            time.sleep(delay)
            return v
        return job

    def bad():
        raise KeyError("nope")

    res = utils.run_and_await_k([ok(1), ok(2), bad], k=2, timeout_after_k=0.2)
    assert res[:2] == [1, 2] and isinstance(res[2], KeyError)
    # stragglers are cut off after timeout_after_k
    t0 = time.time()
    res = utils.run_and_await_k([ok(1), ok(2, delay=3.0)], k=1, timeout_after_k=0.1)
    assert time.time() - t0 < 1.5 and res[0] == 1 and isinstance(res[1], BaseException)
    with pytest.raises(ValueError):
        utils.run_and_await_k([bad, bad, ok(1)], k=2)
    with pytest.raises(TimeoutError):
        utils.run_and_await_k([ok(1, delay=2.0)], k=1, timeout_total=0.1)


def test_countdown_and_await_first():
    ev = utils.CountdownEvent(count_to=2)
    assert not ev.is_set()
    ev.increment(); assert not ev.is_set()
    ev.increment(); assert ev.is_set()
    ev.clear(); assert not ev.is_set()
    e1, e2 = threading.Event(), threading.Event()
    threading.Timer(0.05, e2.set).start()
    assert utils.await_first(e1, e2, k=1, timeout=2)
    fut = utils.run_in_background(lambda a, b: a + b, 1, b=2)
    assert fut.result(timeout=2) == 3


def test_shared_arrays():
    arrays = utils.SharedArrays()
    proto = utils.ArrayProto.from_array(np.zeros((4, 3), dtype=np.float32))
    arr = arrays.create_array("x", proto)
    arr[...] = np.arange(12, dtype=np.float32).reshape(4, 3)
    assert "x" in arrays and len(arrays) == 1
    other = arrays.fork()
    assert np.array_equal(other["x"], arr)
    other["x"][0, 0] = 100
    assert arr[0, 0] == 100  # same memory
    sa = utils.SharedArray.from_array(np.ones(5), shm_manager=arrays.shm_manager)
    arrays["y"] = sa
    assert np.array_equal(arrays["y"], np.ones(5))
    with pytest.raises(ValueError):
        arrays["z"] = np.ones(3)
    del arrays["y"]
    assert "y" not in arrays
    arrays.shm_manager.shutdown()


def test_check_numpy():
    t = torch.arange(3.0, requires_grad=True)
    assert isinstance(utils.check_numpy(t), np.ndarray) and utils.check_numpy([1, 2]).shape == (2,)


def test_lib_alias_exposes_the_reference_public_surface():
    """`import lib` (README snippets / experiment scripts of the reference) resolves every public name of SURVEY.md §1"""
    import lib
    names = ("RemoteExpert GatingFunction TesseractServer TesseractNetwork ExpertBackend TesseractRuntime TaskPool "
             "BatchTensorProto TensorProto ArrayProto Connection PytorchSerializer PickleSerializer nested_compare "
             "nested_flatten nested_pack nested_map SharedArrays SharedArray SharedFuture run_in_background repeated "
             "CountdownEvent await_first run_and_await_k check_numpy DUMMY DUMMY_BATCH_SIZE").split()
    assert [n for n in names if not hasattr(lib, n)] == []
    assert lib.server.TesseractServer is lib.TesseractServer and lib.client.RemoteExpert is lib.RemoteExpert
    assert lib.network.HEARTBEAT_EXPIRATION == 120 and lib.network.UID_DELIMETER == "."
    assert lib.Connection.header_size == 4 and lib.Connection.payload_length_size == 8 and lib.DUMMY_BATCH_SIZE == 3


def test_reference_module_paths_resolve():
    """a user of the reference imports by ITS module paths (lib/utils/threading.py, experiments/convergence/dmoe_emulator.py,
    ...): every one of them must resolve here"""
    import importlib
    import lah_b200  # noqa: F401  (installs the `lib` alias)
    for name in ["lib.utils.connection", "lib.utils.data", "lib.utils.nested", "lib.utils.proto", "lib.utils.serializer",
                 "lib.utils.shared_arrays", "lib.utils.shared_future", "lib.utils.threading", "lib.client.gating_function",
                 "lib.client.remote_expert", "lib.runtime.expert_backend", "lib.runtime.task_pool",
                 "lib.server.connection_handler", "lib.server.network_handler", "lib.network",
                 "lah_b200.experiments.throughput.layers", "lah_b200.experiments.throughput.throughput_server",
                 "lah_b200.experiments.throughput.throughput_client", "lah_b200.experiments.throughput.baseline_throughput",
                 "lah_b200.experiments.throughput.rpc_throughput", "lah_b200.experiments.convergence.dmoe_emulator",
                 "lah_b200.experiments.convergence.faulty_dmoe_emulator"]:
        importlib.import_module(name)
    from lib.utils.threading import run_and_await_k, CountdownEvent  # noqa: F401
    from lah_b200.experiments.convergence.faulty_dmoe_emulator import EmulatedFaultyDMoE, get_non_expert_params  # noqa: F401
    import threading as std_threading
    assert hasattr(std_threading, "Thread")   # the alias module does not shadow the standard library
