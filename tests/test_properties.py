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


@settings(max_examples=80, deadline=None)
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
