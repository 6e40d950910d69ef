"""DHT, in-box index, beam search, GatingFunction (CPU).  Reference: lib/network/__init__.py, lib/client/gating_function.py"""
import itertools
import time

import numpy as np
import pytest
import torch

import lah_b200 as lib
from lah_b200.models import FeedforwardBlock
from lah_b200.network.dht import DHTNode, RoutingTable, sha1, xor_distance


def test_routing_table_native_matches_python_ordering():
    me = sha1(b"me")
    table = RoutingTable(me)
    ids = [sha1(f"node{i}".encode()) for i in range(50)]
    kept = [nid for i, nid in enumerate(ids) if table.add(nid, ("127.0.0.1", 1000 + i))]
    # k-buckets hold at most k=20 contacts: the far half of the id space overflows its bucket
    assert len(table) == len(kept) and 20 < len(kept) < 50 and not table.add(me, ("127.0.0.1", 1))
    target = sha1(b"target")
    got = [nid for nid, _ in table.closest(target, 7)]
    assert got == sorted(kept, key=lambda nid: xor_distance(nid, target))[:7]
    table.remove(got[0])
    assert table.closest(target, 1)[0][0] == got[1]
    assert table.closest(target, 3)[0][1][0] == "127.0.0.1"


@pytest.fixture
def dht_nodes():
    first = lib.TesseractNetwork(port=0, start=True)
    others = [lib.TesseractNetwork(("127.0.0.1", first.port), port=0, start=True) for _ in range(3)]
    yield [first] + others
    for node in [first] + others:
        node.shutdown()


def test_dht_declare_get_first_k_active(dht_nodes):
    a, b, c, d = dht_nodes
    uids = [f"ffn.{i}.{j}" for i in range(3) for j in range(2)]
    b.declare_experts(uids[:4], "10.0.0.1", 1234, wait_timeout=5)
    c.declare_experts(uids[4:], "10.0.0.2", 4321, wait_timeout=5)
    found = d.get_experts(uids + ["ffn.9.9"])
    assert [e.uid if e else None for e in found] == uids + [None]
    assert (found[0].host, found[0].port) == ("10.0.0.1", 1234) and (found[5].host, found[5].port) == ("10.0.0.2", 4321)
    # prefixes: every uid prefix gets a heartbeat
    assert a.first_k_active(["ffn.7", "ffn.2", "ffn.0", "ffn.1"], k=2) == ["ffn.2", "ffn.0"]
    assert a.first_k_active(["nope", "ffn.1.1", "ffn.1.5"], k=3) == ["ffn.1.1"]
    # expiration: a heartbeat older than the limit counts as dead
    time.sleep(0.3)
    assert d.get_experts(uids[:1], heartbeat_expiration=0.1) == [None]
    assert a.first_k_active(["ffn.0"], k=1, heartbeat_expiration=0.1) == []
    assert lib.TesseractNetwork.make_key("expert", "x") == "expert::x" and lib.TesseractNetwork.UID_DELIMETER == "."


def test_inbox_network_same_api():
    net = lib.InBoxNetwork()
    uids = [f"e.{i}.{j}" for i in range(4) for j in range(4) if (i + j) % 3]
    net.declare_experts(uids, "127.0.0.1", 8080, owner=1)
    got = net.get_experts(["e.0.1", "e.0.0", "e.3.3"])
    assert [g.uid if g else None for g in got] == ["e.0.1", None, None]
    assert net.first_k_active(["e.9", "e.1", "e.0"], k=2) == ["e.1", "e.0"]
    mask = net.alive_mask((4, 4), "e")
    expect = torch.tensor([1 if (i + j) % 3 else 0 for i in range(4) for j in range(4)], dtype=torch.uint8)
    assert torch.equal(mask, expect)
    time.sleep(0.15)
    assert net.get_experts(["e.0.1"], heartbeat_expiration=0.05) == [None]


def brute_force_topk(scores, alive_uids, prefix, k):
    grid = [s.shape[1] for s in scores]
    out = []
    for b in range(scores[0].shape[0]):
        cands = []
        for coords in itertools.product(*(range(g) for g in grid)):
            uid = ".".join([prefix] + [str(c) for c in coords])
            if uid in alive_uids:
                cands.append((sum(float(scores[d][b, c]) for d, c in enumerate(coords)), uid))
        out.append([uid for _, uid in sorted(cands, key=lambda t: -t[0])[:k]])
    return out


@pytest.mark.parametrize("grid", [(4, 5), (3, 3, 3)])
def test_beam_search_equals_brute_force_over_alive_experts(grid):
    torch.manual_seed(0)
    net = lib.InBoxNetwork()
    all_uids = [".".join(["ex"] + [str(c) for c in coords]) for coords in itertools.product(*(range(g) for g in grid))]
    rng = np.random.RandomState(0)
    alive = [uid for uid in all_uids if rng.rand() > 0.4]
    net.declare_experts(alive, "127.0.0.1", 1)
    gate = lib.GatingFunction(in_features=8, grid_size=grid, network=net, k_best=3, uid_prefix="ex")
    scores = [torch.randn(6, g) for g in grid]
    chosen = gate.beam_search(scores, 3)
    assert [[e.uid for e in row] for row in chosen] == brute_force_topk(scores, set(alive), "ex", 3)
    # fewer alive experts than k_best: no crash (the reference dies on None + '.')
    sparse = lib.InBoxNetwork()
    sparse.declare_experts(alive[:2], "127.0.0.1", 1)
    gate2 = lib.GatingFunction(in_features=8, grid_size=grid, network=sparse, k_best=3, uid_prefix="ex")
    assert all(1 <= len(row) <= 2 for row in gate2.beam_search(scores, 3))
    gate.close(), gate2.close()


def test_score_experts_gradient_matches_dense_reference():
    net = lib.InBoxNetwork()
    uids = [f"g.{i}.{j}" for i in range(3) for j in range(4)]
    net.declare_experts(uids, "127.0.0.1", 1)
    gate = lib.GatingFunction(in_features=8, grid_size=(3, 4), network=net, k_best=2, uid_prefix="g")
    s0, s1 = torch.randn(5, 3, requires_grad=True), torch.randn(5, 4, requires_grad=True)
    chosen = gate.beam_search([s0, s1], 2)
    logits = gate._score_experts([s0, s1], chosen)
    total = sum(v for row in logits for v in row.values())
    total.backward()
    dense = (s0.detach()[:, :, None] + s1.detach()[:, None, :])
    g0, g1 = torch.zeros(5, 3), torch.zeros(5, 4)
    for b, row in enumerate(chosen):
        for e in row:
            i, j = map(int, e.uid.split(".")[1:])
            assert torch.isclose(logits[b][e].detach(), dense[b, i, j])
            g0[b, i] += 1
            g1[b, j] += 1
    assert torch.equal(s0.grad, g0) and torch.equal(s1.grad, g1)
    gate.close()


def _make_server(uids, hid, network):
    experts = {}
    for uid in uids:
        block = FeedforwardBlock(hid)
        experts[uid] = lib.ExpertBackend(name=uid, expert=block, opt=torch.optim.Adam(block.parameters(), amsgrad=True),
                                         args_schema=(lib.BatchTensorProto(hid),), outputs_schema=lib.BatchTensorProto(hid),
                                         max_batch_size=256)
    return lib.TesseractServer(network, experts, port=0, conn_handler_processes=4, update_period=1).run_in_background()


def test_gating_function_end_to_end_and_fault_tolerance():
    torch.manual_seed(1)
    net = lib.InBoxNetwork()
    uids = [f"expert.{i}.{j}" for i in range(2) for j in range(3)]
    server = _make_server(uids, 16, net)
    time.sleep(0.2)
    gate = lib.GatingFunction(in_features=16, grid_size=(2, 3), network=net, k_best=3, k_min=1,
                              timeout_after_k_min=2.0, uid_prefix="expert")
    x = torch.randn(6, 16, requires_grad=True)
    out = gate(x)
    # dense reference with the same experts
    scores = gate.proj(x.detach()).split_with_sizes((2, 3), dim=-1)
    full = (scores[0][:, :, None] + scores[1][:, None, :]).flatten(1)
    top_v, top_i = full.topk(3, dim=-1)
    w = torch.softmax(top_v, -1)
    ref = torch.zeros(6, 16)
    for b in range(6):
        for j in range(3):
            uid = f"expert.{top_i[b, j] // 3}.{top_i[b, j] % 3}"
            ref[b] += w[b, j] * server.experts[uid].expert(x.detach()[b: b + 1])[0]
    assert torch.allclose(out, ref, atol=1e-5)
    out.sum().backward()
    assert gate.proj.weight.grad.abs().sum() > 0 and x.grad.shape == x.shape
    assert sum(be.update_count for be in server.experts.values()) >= 1  # experts trained server-side

    # an expert that is declared alive but unreachable is dropped; weights renormalise over the responders
    net.declare_experts(["expert.1.1"], "127.0.0.1", 1)  # nothing listens on port 1
    x2 = torch.randn(4, 16)
    out2 = gate(x2)
    assert out2.shape == (4, 16) and torch.isfinite(out2).all()
    gate.k_min = 3
    with pytest.raises(ValueError):
        gate(x2)
    gate.close()
    server.shutdown()


def test_dht_rejects_malformed_and_executable_messages():
    """the DHT envelope is msgpack with validated fields — nothing from the network is ever unpickled, and ids that
    are not 20 bytes never reach the native routing table"""
    import os
    import pickle
    import msgpack
    from lah_b200.network import dht, _loads_value
    good = dict(t="q", id=os.urandom(8), m="find_node", sender=os.urandom(20), target=os.urandom(20))
    assert dht._valid_message(good)
    for bad in (dict(good, sender=b"short"), dict(good, target=b"x" * 19), dict(good, m="exec"), dict(good, id=b"1"),
                dict(good, t="z"), dict(good, nodes=[[os.urandom(20), ["not an ip", 5]]]),
                dict(good, nodes=[[os.urandom(3), ["127.0.0.1", 5]]]), {k: v for k, v in good.items() if k != "target"}):
        assert not dht._valid_message(bad), bad
    node = dht.DHTNode()
    proto = dht._Protocol(node)

    class Boom:
        def __reduce__(self):
            return (os.system, ("echo pwned > /tmp/lah_dht_pwned",))

    proto.datagram_received(pickle.dumps(Boom()), ("127.0.0.1", 1))          # a pickle is just an invalid msgpack message
    proto.datagram_received(msgpack.packb(dict(good, target=b"abc"), use_bin_type=True), ("127.0.0.1", 1))
    assert not os.path.exists("/tmp/lah_dht_pwned") and len(node.table) == 0
    table = dht.RoutingTable(os.urandom(20))
    assert not table.add(b"short-id", ("127.0.0.1", 1))
    import pytest
    with pytest.raises(ValueError):
        table.closest(b"way-too-short")
    # stored values: the reference's pickled records load, anything naming other globals does not
    import datetime
    rec = ((("127.0.0.1", 8080)), datetime.datetime.now())
    assert _loads_value(pickle.dumps(rec))[0] == ("127.0.0.1", 8080)
    with pytest.raises(pickle.UnpicklingError):
        _loads_value(pickle.dumps(Boom()))
