"""Property-based tests (hypothesis) of the host-side building blocks: the raw-tensor wire format of the TCP path, the
canonical flattening order used on both ends of the wire (reference: lib/utils/nested.py) and the hot-expert shadow plan that
`layout_exchange_kernel` implements on the device."""
import struct

import pytest
import torch

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

import lah_b200  # noqa: F401,E402
from lah_b200.parallel import balance  # noqa: E402
from lah_b200.utils import nested, tensor_wire  # noqa: E402

DTYPES = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.int16, torch.int8,
          torch.uint8, torch.bool]


@st.composite
def tensors(draw):
    dtype = draw(st.sampled_from(DTYPES))
    shape = draw(st.lists(st.integers(0, 5), min_size=0, max_size=4))
    if dtype == torch.bool:
        return torch.rand(shape) > 0.5
    if dtype.is_floating_point:
        return torch.randn(shape).to(dtype)
    return torch.randint(-100 if dtype != torch.uint8 else 0, 100, shape).to(dtype)


@settings(max_examples=60, deadline=None)
@given(uid=st.text(alphabet="abcdefghij.0123456789", min_size=0, max_size=40), ts=st.lists(tensors(), min_size=0, max_size=5))
def test_tensor_wire_roundtrip(uid, ts):
    parts, total = tensor_wire.encode(uid, ts)
    blob = bytearray().join(bytes(p) for p in parts)
    assert len(blob) == total
    uid2, out = tensor_wire.decode(blob)
    assert uid2 == uid and len(out) == len(ts)
    for a, b in zip(ts, out):
        assert a.dtype == b.dtype and tuple(a.shape) == tuple(b.shape)
        assert torch.equal(a, b) or (a.dtype.is_floating_point and torch.equal(a.view(torch.uint8) if a.numel() == 0 else a, b))


@settings(max_examples=12, deadline=None)
@given(ts=st.lists(tensors(), min_size=1, max_size=3), cut=st.integers(0, 10 ** 6), flip=st.integers(0, 10 ** 6),
       val=st.integers(0, 255))
def test_tensor_wire_rejects_damaged_frames_without_crashing(ts, cut, flip, val):
    """every field of a frame is attacker-controlled: truncation or a flipped byte must end in ValueError (or in a decoded
    frame whose tensors lie inside the buffer), never in an out-of-bounds read or another exception type"""
    parts, total = tensor_wire.encode("e.1", ts)
    blob = bytearray().join(bytes(p) for p in parts)
    damaged = bytearray(blob[: cut % (total + 1)])
    if damaged:
        damaged[flip % len(damaged)] = val
    try:
        _, out = tensor_wire.decode(damaged)
    except (ValueError, struct.error, UnicodeDecodeError):
        return
    assert sum(t.numel() * t.element_size() for t in out) <= len(damaged)


leaves = st.integers(-5, 5)
trees = st.recursive(leaves, lambda kids: st.one_of(st.lists(kids, max_size=3), st.tuples(kids, kids),
                                                    st.dictionaries(st.sampled_from("abcd"), kids, max_size=3)), max_leaves=12)


@settings(max_examples=100, deadline=None)
@given(tree=trees)
def test_nested_flatten_pack_roundtrip(tree):
    flat = list(nested.nested_flatten(tree))
    assert nested.nested_compare(nested.nested_pack(flat, tree), tree)
    assert list(nested.nested_flatten(nested.nested_pack(flat, tree))) == flat
    doubled = nested.nested_map(lambda v: 2 * v, tree)
    assert list(nested.nested_flatten(doubled)) == [2 * v for v in flat]


@settings(max_examples=100, deadline=None)
@given(world=st.integers(1, 8), e_loc=st.integers(1, 6), max_shadow=st.integers(0, 8), data=st.data())
def test_shadow_plan_invariants(world, e_loc, max_shadow, data):
    E = world * e_loc
    counts = [[data.draw(st.integers(0, 50)) for _ in range(E)] for _ in range(world)]
    before = balance.rank_loads(counts, e_loc)
    shadowed, loads = balance.shadow_plan(counts, e_loc, max_shadow, tol=1.05)
    assert len(shadowed) <= max_shadow and len(set(shadowed)) == len(shadowed)
    assert sum(loads) == sum(before) == sum(map(sum, counts))            # shadowing moves rows, it never drops any
    assert loads == balance.rank_loads(counts, e_loc, shadowed)          # incremental bookkeeping == recomputation
    # NOT an invariant: max(loads) <= max(before).  Shadowing hands an expert's rows back to their senders; a sender that is
    # itself nearly the most loaded rank and sent most of those rows ends up above the old maximum (hypothesis finds such
    # count matrices at once).  Real routing sends every expert rows from all ranks alike, so the owner sheds (1 - 1/world)
    # of the expert and every sender gains 1/world of it.  What does hold: the rank that was relieved never gains rows.
    for e in shadowed:
        owner = e // e_loc
        assert balance.rank_loads(counts, e_loc, [e])[owner] <= before[owner]
    if world == 1:
        assert shadowed == []


# ------------------------------------------------------------------------------------------------ DHT datagrams (untrusted)
msgpack_values = st.recursive(
    st.one_of(st.none(), st.booleans(), st.integers(-2 ** 31, 2 ** 31), st.binary(max_size=24), st.text(max_size=12)),
    lambda kids: st.one_of(st.lists(kids, max_size=4), st.dictionaries(st.text(max_size=8), kids, max_size=5)), max_leaves=10)


@settings(max_examples=150, deadline=None)
@given(raw=st.binary(max_size=200), obj=msgpack_values, field=st.sampled_from(["t", "id", "m", "sender", "target", "key", "nodes",
                                                                              "value", "found"]))
def test_dht_datagram_handler_survives_arbitrary_input(raw, obj, field):
    """raw bytes, arbitrary msgpack objects and well-formed queries with ONE field replaced by an arbitrary object: the handler
    must neither raise nor learn a malformed node id"""
    import os
    import msgpack
    from lah_b200.network import dht
    node = dht.DHTNode()
    proto = dht._Protocol(node)
    good = dict(t="q", id=os.urandom(8), m="find_node", sender=os.urandom(20), target=os.urandom(20))
    proto.datagram_received(raw, ("127.0.0.1", 9))
    proto.datagram_received(msgpack.packb(obj, use_bin_type=True), ("127.0.0.1", 9))
    mutated = dict(good)
    mutated[field] = obj
    try:
        proto.datagram_received(msgpack.packb(mutated, use_bin_type=True), ("127.0.0.1", 9))
    except AttributeError as e:   # a valid query is answered through the transport, which this bare protocol does not have
        assert "transport" in str(e) or "sendto" in str(e), e
    for nid in node.table.ids() if hasattr(node.table, "ids") else []:
        assert isinstance(nid, bytes) and len(nid) == 20


# ------------------------------------------------------------------------------------------------ gate oracle, Adam oracle
@settings(max_examples=60, deadline=None)
@given(grid=st.lists(st.integers(1, 5), min_size=1, max_size=3), batch=st.integers(1, 6), k=st.integers(1, 4), data=st.data())
def test_gate_oracle_selects_the_best_alive_experts(grid, batch, k, data):
    from lah_b200.ops import kernels as K
    E = 1
    for g in grid:
        E *= g
    k = min(k, E)
    logits = torch.randn(batch, sum(grid), generator=torch.Generator().manual_seed(data.draw(st.integers(0, 10 ** 6))))
    alive = torch.tensor([data.draw(st.booleans()) for _ in range(E)])
    idx, w = K.gate_topk_ref(logits, grid, k, alive=alive)
    scores = K.product_key_scores(logits, grid)
    n_alive = int(alive.sum())
    for b in range(batch):
        chosen = [int(i) for i in idx[b] if i >= 0]
        assert len(chosen) == min(k, n_alive) and len(set(chosen)) == len(chosen)
        assert all(bool(alive[i]) for i in chosen)
        if chosen:
            worst = min(float(scores[b, i]) for i in chosen)
            others = [float(scores[b, i]) for i in range(E) if alive[i] and i not in chosen]
            assert all(o <= worst + 1e-6 for o in others)                     # nothing alive and better was left out
            assert abs(float(w[b].sum()) - 1.0) < 1e-5 and bool((w[b][idx[b] < 0] == 0).all())
        else:
            assert float(w[b].abs().sum()) == 0.0


@settings(max_examples=25, deadline=None)
@given(G=st.integers(1, 4), sizes=st.lists(st.integers(1, 6), min_size=1, max_size=3), steps=st.integers(1, 4),
       amsgrad=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_adam_oracle_equals_torch_adam_per_group(G, sizes, steps, amsgrad, seed):
    """csrc/adam.cu's PyTorch twin (flat segments [G, size], per-group step counts, inactive groups untouched) == one
    torch.optim.Adam per expert (reference: one optimizer per ExpertBackend, lib/runtime/expert_backend.py:95-97)"""
    from lah_b200.ops import kernels as K
    gen = torch.Generator().manual_seed(seed)
    sizes = [4 * s for s in sizes]
    total = G * sum(sizes)
    p = torch.randn(total, generator=gen)
    m, v, vmax = torch.zeros(total), torch.zeros(total), torch.zeros(total)
    step = torch.zeros(G, dtype=torch.int32)

    def views(flat):   # per group: list of its segment slices
        out, off = [[] for _ in range(G)], 0
        for s in sizes:
            for g in range(G):
                out[g].append(flat[off + g * s: off + (g + 1) * s])
            off += s * G
        return out

    params = [[t.clone().requires_grad_(True) for t in group] for group in views(p)]
    opts = [torch.optim.Adam(group, lr=1e-2, amsgrad=amsgrad) for group in params]
    for _ in range(steps):
        g = torch.randn(total, generator=gen)
        rows = torch.randint(0, 2, (G,), generator=gen).to(torch.int32)
        step += (rows > 0).to(torch.int32)
        K.adam_step_ref(p, g.clone(), m, v, vmax, sizes, G, step=step, group_rows=rows, lr=1e-2, amsgrad=amsgrad)
        for gi, (group, gviews) in enumerate(zip(params, views(g))):
            if rows[gi] > 0:
                for t, gv in zip(group, gviews):
                    t.grad = gv.clone()
                opts[gi].step()
    for group, pviews in zip(params, views(p)):
        for t, pv in zip(group, pviews):
            assert torch.allclose(t.detach(), pv, atol=2e-6, rtol=1e-5)


# ------------------------------------------------------------------------------------------------ server-side batching, k-of-n
@settings(max_examples=40, deadline=None)
@given(sizes=st.lists(st.integers(1, 6), min_size=1, max_size=12), max_batch=st.integers(1, 16))
def test_task_pool_batches_preserve_order_and_rows(sizes, max_batch):
    """dynamic batching (reference lib/runtime/task_pool.py:105-125): requests are served in submission order, a batch is a
    prefix of the queue that stops with the first task that reaches max_batch_size (overshoot by at most one task), every row
    comes back to the future of the request that sent it"""
    from lah_b200.runtime import TaskPool
    schema = (lah_b200.BatchTensorProto(2),)
    pool = TaskPool(lambda x: (x + 1,), schema, schema, max_batch_size=max_batch, pool_size=len(sizes) + 1)
    xs = [torch.full((n, 2), float(i)) for i, n in enumerate(sizes)]
    futures = [pool.submit_task(x) for x in xs]
    served = 0
    while served < len(xs):
        index, (batch,) = pool.load_batch_to_runtime(timeout=1)
        rows, taken = 0, 0
        while served + taken < len(xs) and rows < max_batch:      # greedy prefix: add tasks until the size is reached
            rows += sizes[served + taken]
            taken += 1
        assert batch.shape[0] == rows and torch.equal(batch, torch.cat(xs[served: served + taken]))
        pool.send_outputs_from_runtime(index, pool.process_func(batch))
        served += taken
    for fut, x in zip(futures, xs):
        assert torch.equal(fut.result(timeout=1)[0], x + 1)
    assert pool.empty


@settings(max_examples=40, deadline=None)
@given(outcomes=st.lists(st.booleans(), min_size=1, max_size=8), data=st.data())
def test_run_and_await_k_returns_every_outcome_or_raises(outcomes, data):
    """reference lib/utils/threading.py:76-125: with k successes available every job's outcome is reported (value or
    exception); with fewer than k possible successes the call raises instead of hanging"""
    from lah_b200.utils.threads import run_and_await_k
    k = data.draw(st.integers(0, len(outcomes)))

    def job(i, ok):
        def run():
            if not ok:
                raise KeyError(i)
            return i
        return run

    jobs = [job(i, ok) for i, ok in enumerate(outcomes)]
    if sum(outcomes) >= k:
        res = run_and_await_k(jobs, k, timeout_after_k=1.0, timeout_total=5.0)
        assert len(res) == len(jobs)
        for i, (ok, r) in enumerate(zip(outcomes, res)):
            assert (r == i) if ok else isinstance(r, (KeyError, TimeoutError))
    else:
        with pytest.raises(ValueError):
            run_and_await_k(jobs, k, timeout_after_k=1.0, timeout_total=5.0)


# ------------------------------------------------------------------------------------------------ trainer configurations
@settings(max_examples=12, deadline=None)
@given(grid=st.sampled_from([(4,), (2, 2), (2, 3), (8,)]), k=st.integers(1, 4), layers=st.integers(1, 2),
       micro=st.sampled_from([1, 2, 4]), stale=st.integers(0, 2), every=st.sampled_from([(0, 0), (10 ** 6, 2), (24, 0)]),
       gate=st.sampled_from(["emulator", "product_key"]), fail=st.sampled_from([0.0, 0.3]))
def test_cpu_trainer_any_configuration_steps_and_resumes(grid, k, layers, micro, stale, every, gate, fail):
    """every combination of the asynchronous-training knobs (micro-batched trainers, stale trainer gradients, lazily stepped
    experts, failure injection, both gates) trains on the CPU oracle path with finite losses, and a checkpoint taken mid-run
    resumes to the same losses"""
    from lah_b200.parallel import engine as E
    from lah_b200.parallel.trainer import DMoETrainer
    n_experts = 1
    for g in grid:
        n_experts *= g
    if gate == "emulator":
        grid = (n_experts,)   # the emulator gate scores the experts densely: 1-D grid by construction
    cfg = E.DMoEConfig(hidden=16, grid_size=grid, k=min(k, n_experts), num_layers=layers, in_features=8, tokens_per_rank=16,
                       lr=1e-3, trainer_microbatches=micro, trainer_staleness=stale, update_every_inputs=every[0],
                       update_every_steps=every[1], gate_mode=gate, failure_rate=fail, seed=3)
    torch.manual_seed(1)
    x, y = torch.randn(16, 8), torch.randint(0, 10, (16,))
    threads = torch.get_num_threads()
    torch.set_num_threads(1)   # tiny tensors: intra-op threads only fight the background threads earlier tests left behind
    try:
        trainer = DMoETrainer(cfg)
        losses = [trainer.train_step(x, y) for _ in range(3)]
        assert all(l == l and abs(l) < 1e4 for l in losses), losses
        if fail == 0.0:   # (failure masks come from the global RNG on the oracle path: only deterministic runs are compared)
            clone = DMoETrainer(cfg)
            clone.load_state_dict(trainer.state_dict())
            for _ in range(2):
                assert abs(trainer.train_step(x, y) - clone.train_step(x, y)) < 1e-5
    finally:
        torch.set_num_threads(threads)


# ------------------------------------------------------------------------------------------------ expert index + beam search
@settings(max_examples=40, deadline=None)
@given(grid=st.lists(st.integers(1, 4), min_size=1, max_size=3), k=st.integers(1, 5), batch=st.integers(1, 4), data=st.data())
def test_index_and_beam_search_agree_with_brute_force(grid, k, batch, data):
    """in-box expert index (declare_experts / get_experts / first_k_active / alive_mask) against plain Python sets, and
    GatingFunction.beam_search against brute-force top-k over the ALIVE experts of the product grid (reference:
    lib/client/gating_function.py:68-123, lib/network/__init__.py:35-129), including grids with fewer alive experts than k"""
    import itertools
    import lah_b200 as lib
    all_uids = [".".join(["px"] + [str(c) for c in coords]) for coords in itertools.product(*(range(g) for g in grid))]
    alive = [uid for uid in all_uids if data.draw(st.booleans())]
    net = lib.InBoxNetwork()
    net.declare_experts(alive, "127.0.0.1", 7)
    assert [e is not None for e in net.get_experts(all_uids)] == [uid in set(alive) for uid in all_uids]
    assert net.alive_mask(grid, "px").tolist() == [int(uid in set(alive)) for uid in all_uids]
    prefixes = sorted({".".join(uid.split(".")[: j + 1]) for uid in all_uids for j in range(1, len(grid) + 1)})
    alive_prefixes = {".".join(uid.split(".")[: j + 1]) for uid in alive for j in range(len(grid) + 1)}
    assert net.first_k_active(prefixes, k) == [p for p in prefixes if p in alive_prefixes][:k]
    gate = lib.GatingFunction(in_features=4, grid_size=grid, network=net, k_best=k, uid_prefix="px")
    seed = data.draw(st.integers(0, 10 ** 6))
    gen = torch.Generator().manual_seed(seed)
    scores = [torch.randn(batch, g, generator=gen) for g in grid]
    chosen = gate.beam_search(scores, k)
    for b in range(batch):
        total = {uid: sum(float(scores[d][b, int(c)]) for d, c in enumerate(uid.split(".")[1:])) for uid in alive}
        best = sorted(total, key=lambda u: -total[u])[:k]
        got = [e.uid for e in chosen[b]]
        assert set(got) <= set(alive) and len(set(got)) == len(got) and len(got) <= k and (len(got) >= 1 or not alive)
        got_scores = [total[u] for u in got]
        assert all(a >= b - 1e-9 for a, b in zip(got_scores, got_scores[1:]))          # best first
        if len(alive) == len(all_uids):
            # additive scores + nothing pruned by liveness: a width-k beam is EXACT (a top-k expert's prefix is a top-k prefix)
            assert [round(v, 5) for v in got_scores] == [round(total[u], 5) for u in best]  # (ties may swap uids)
        # with holes in the grid the beam is a heuristic, exactly like the reference's: a live prefix can win a level on its
        # own score and then have only poor descendants (hypothesis: grid [3, 1, 4], k = 1, two far-apart alive experts).
        # The fused gate (gate_topk_kernel / gate_topk_ref) scores all alive experts and IS exact — tested above.
    gate.close()


# ------------------------------------------------------------------------------------------------ TCP framing
@settings(max_examples=25, deadline=None)
@given(header=st.text(alphabet="abcdefghijklmnopqrstuvwxyz_", min_size=4, max_size=4), payload=st.binary(max_size=5000),
       chunk=st.integers(1, 700), eager=st.sampled_from([16, 1000, 64 << 20]))
def test_connection_framing_with_arbitrary_fragmentation(header, payload, chunk, eager):
    """header(4) | length(8, big endian) | payload (reference lib/utils/connection.py): the receiver reassembles a message that
    arrives in arbitrary fragments, whether the buffer was pre-sized or grown as the bytes arrived"""
    import socket
    import threading
    from lah_b200.utils.connection import Connection
    a, b = socket.socketpair()
    rx = Connection(b, ("local", 0))
    rx.eager_alloc = eager
    wire = header.encode() + len(payload).to_bytes(8, "big") + payload

    def feed():
        for i in range(0, len(wire), chunk):
            a.sendall(wire[i: i + chunk])

    t = threading.Thread(target=feed)
    t.start()
    try:
        assert rx.recv_message() == (header, payload)
    finally:
        t.join()
        a.close()
        rx.close()


def test_connection_does_not_allocate_what_a_peer_merely_announces():
    import socket
    from lah_b200.utils.connection import Connection
    a, b = socket.socketpair()
    rx = Connection(b, ("local", 0))
    a.sendall(b"fwd_" + (3 << 30).to_bytes(8, "big") + b"x" * 10)   # announces 3 GiB, sends 10 bytes, hangs up
    a.close()
    assert rx.recv_header() == "fwd_"
    with pytest.raises(RuntimeError):
        rx.recv_raw()
    rx.close()
    a2, b2 = socket.socketpair()
    rx2 = Connection(b2, ("local", 0))
    a2.sendall(b"fwd_" + (1 << 40).to_bytes(8, "big"))             # beyond max_payload: refused before any allocation
    assert rx2.recv_header() == "fwd_"
    with pytest.raises(ValueError):
        rx2.recv_raw()
    a2.close(), rx2.close()


# ------------------------------------------------------------------------------------------------ live server under fuzz
@pytest.fixture(scope="module")
def fuzz_server():
    from lah_b200.models import FeedforwardBlock
    expert = FeedforwardBlock(16)
    backend = lah_b200.ExpertBackend(name="e", expert=expert, opt=torch.optim.Adam(expert.parameters(), lr=1e-3),
                                     args_schema=(lah_b200.BatchTensorProto(16),), outputs_schema=lah_b200.BatchTensorProto(16),
                                     max_batch_size=64)
    srv = lah_b200.TesseractServer(None, {"e": backend}, port=0, conn_handler_processes=1)   # ONE acceptor
    srv.run_in_background()
    yield srv
    srv.shutdown()


pickled = st.recursive(st.one_of(st.none(), st.integers(-5, 5), st.text(max_size=5), st.binary(max_size=8)),
                       lambda kids: st.one_of(st.lists(kids, max_size=3), st.tuples(kids, kids)), max_leaves=6)


@settings(max_examples=25, deadline=None, suppress_health_check=list(hypothesis.HealthCheck))
@given(header=st.sampled_from(["fwd_", "bwd_", "info", "fwdT", "bwdT", "zzzz", "rest"]), body=st.binary(max_size=300),
       obj=pickled, lie=st.integers(0, 400))
def test_server_survives_arbitrary_requests(fuzz_server, header, body, obj, lie):
    """raw garbage, a lying length prefix and well-formed pickles of the wrong shape, on every request type: each costs its
    author an error reply or the connection; the single acceptor and the runtime keep serving a correct client"""
    import socket
    from lah_b200.utils import Connection, PytorchSerializer
    port = fuzz_server.port

    def raw(payload: bytes):
        with socket.create_connection(("127.0.0.1", port), timeout=5) as s:
            s.sendall(payload)
            s.settimeout(0.3)
            try:
                s.recv(1 << 16)
            except (socket.timeout, ConnectionResetError, BrokenPipeError):
                pass

    raw(header.encode() + len(body).to_bytes(8, "big") + body)
    raw(header.encode() + lie.to_bytes(8, "big") + body[: max(0, lie - 1)])            # announces more than it sends
    with Connection.create("127.0.0.1", port, timeout=5) as c:
        c.send_raw(header, PytorchSerializer.dumps(("e", obj)))
        c.conn.settimeout(2)
        try:
            reply, _ = c.recv_message()
            assert reply in ("err_", "rest", "resT")
        except (socket.timeout, RuntimeError, ConnectionResetError):
            pass
    remote = lah_b200.RemoteExpert("e", "127.0.0.1", port, timeout=10)
    x = torch.randn(3, 16)
    assert torch.allclose(remote(x), fuzz_server.experts["e"].expert(x), atol=1e-5)


def test_expert_index_grows_and_keeps_every_entry():
    """native open-addressing index (csrc/host_runtime.cpp): 20,000 experts + their prefixes force several rehashes; every
    declared uid stays retrievable with its (owner, slot), absent uids stay absent, a re-declaration overwrites in place"""
    import lah_b200 as lib
    net = lib.InBoxNetwork()
    uids = [f"big.{i // 200}.{i % 200}" for i in range(20000)]
    net.declare_experts(uids, "127.0.0.1", 1, owner=3, slots=list(range(20000)))
    for i in (0, 1, 199, 200, 7777, 19999):
        assert net._fresh("expert", uids[i], 60.0) == (3, i)
    assert all(e is not None for e in net.get_experts(uids[::97]))
    assert net.get_experts(["big.100.200", "big.9999.0", "bigger.0.0"]) == [None, None, None]
    assert net.first_k_active(["big.500", "big.42", "big.99"], 2) == ["big.42", "big.99"]
    net.declare_experts([uids[5]], "127.0.0.1", 1, owner=1, slots=[9])
    assert net._fresh("expert", uids[5], 60.0) == (1, 9)
