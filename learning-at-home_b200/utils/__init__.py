from .connection import Connection
from .data import check_numpy, DUMMY
from .nested import nested_compare, nested_flatten, nested_pack, nested_map, is_namedtuple
from .proto import ProtoBase, ArrayProto, TensorProto, BatchTensorProto, DUMMY_BATCH_SIZE
from .serializer import PickleSerializer, JoblibSerializer, PytorchSerializer
from .shared_arrays import SharedArrays, SharedArray
from .shared_future import SharedFuture
from .threads import (run_in_background, repeated, add_event_callback, CountdownEvent, await_first, run_and_await_k)

__all__ = [
    "Connection", "check_numpy", "DUMMY", "nested_compare", "nested_flatten", "nested_pack", "nested_map",
    "is_namedtuple", "ProtoBase", "ArrayProto", "TensorProto", "BatchTensorProto", "DUMMY_BATCH_SIZE",
    "PickleSerializer", "JoblibSerializer", "PytorchSerializer", "SharedArrays", "SharedArray", "SharedFuture",
    "run_in_background", "repeated", "add_event_callback", "CountdownEvent", "await_first", "run_and_await_k",
]
