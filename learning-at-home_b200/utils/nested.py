"""
Helpers for nested containers (tuples / lists / namedtuples / dicts whose leaves are tensors or schemas).

They define the canonical flattening order used on both ends of the wire: sequences in order, dict values in
SORTED-KEY order (API parity with /root/reference/lib/utils/nested.py:4-97).
"""
from typing import Any, Callable, Iterable, Iterator


def is_namedtuple(obj) -> bool:
    """True for instances of classes produced by collections.namedtuple / typing.NamedTuple."""
    cls = type(obj)
    fields = getattr(cls, "_fields", None)
    return (isinstance(obj, tuple) and cls.__bases__ == (tuple,) and isinstance(fields, tuple)
            and all(isinstance(name, str) for name in fields))


def _children(node):
    """ordered children of a container node, or None for a leaf"""
    if isinstance(node, dict):
        return [node[key] for key in sorted(node)]
    if isinstance(node, (list, tuple)):
        return list(node)
    return None


def nested_compare(t, u) -> bool:
    """Whether two nested structures have the same shape (container types, lengths, dict keys). Leaves always match."""
    stack = [(t, u)]
    while stack:
        a, b = stack.pop()
        if isinstance(a, (list, tuple)):
            if not isinstance(b, type(a)) or len(a) != len(b):
                return False
            stack.extend(zip(a, b))
        elif isinstance(a, dict):
            if not isinstance(b, dict) or a.keys() != b.keys():
                return False
            stack.extend((a[key], b[key]) for key in a)
        # a is a leaf: anything goes (mirrors the reference, which does not inspect u when t is a leaf)
    return True


def nested_flatten(t) -> Iterator[Any]:
    """Depth-first iterator over the leaves of a nested structure."""
    kids = _children(t)
    if kids is None:
        yield t
        return
    for kid in kids:
        yield from nested_flatten(kid)


def nested_pack(flat: Iterable[Any], structure):
    """Inverse of nested_flatten: pour the leaves of :flat: into a container shaped like :structure:."""
    leaves = iter(flat)

    def build(template):
        if is_namedtuple(template):
            return type(template)(*[build(item) for item in template])
        if isinstance(template, (list, tuple)):
            return type(template)(build(item) for item in template)
        if isinstance(template, dict):
            return {key: build(template[key]) for key in sorted(template)}
        return next(leaves)

    return build(structure)


def nested_map(fn: Callable, *structures):
    """Apply fn leaf-wise over one or more identically shaped structures; result is shaped like the first one."""
    if not structures:
        raise ValueError("nested_map expects a function and at least one structure")
    head = structures[0]
    for other in structures[1:]:
        if not nested_compare(head, other):
            raise ValueError(f"Nested structure of {head!r} and {other!r} differs")
    return nested_pack(map(fn, *map(nested_flatten, structures)), head)
