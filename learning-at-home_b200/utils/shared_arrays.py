"""
Numpy arrays in POSIX shared memory, addressable by key from several processes
(parity: /root/reference/lib/utils/shared_arrays.py:12-100).

Fixes relative to the reference: segments are always allocated through the owning
SharedMemoryManager (the reference's ``SharedArray.from_array`` leaks bare segments), and attached segments are cached
per process instead of being re-mmapped on every ``__getitem__``.
"""
import multiprocessing as mp
import multiprocessing.managers
import multiprocessing.shared_memory

import numpy as np

from .proto import ArrayProto


class SharedArray(np.ndarray):
    """ndarray view over a SharedMemory block; keeps the block alive for as long as any view of it exists."""

    def __new__(cls, proto: ArrayProto, shared_memory: mp.shared_memory.SharedMemory, offset: int = 0):
        obj = super().__new__(cls, proto.shape, proto.dtype, shared_memory.buf, offset, proto.strides, proto.order)
        obj.shared_memory = shared_memory
        return obj

    def __array_finalize__(self, parent):
        if parent is not None:  # views / slices inherit the owner so the mapping cannot be closed under them
            self.shared_memory = getattr(parent, "shared_memory", None)

    def __array_wrap__(self, out_arr, context=None, return_scalar=False):
        return np.asarray(out_arr)  # results of out-of-place ops are ordinary arrays

    @classmethod
    def from_array(cls, arr: np.ndarray, shared_memory: mp.shared_memory.SharedMemory = None, shm_manager=None):
        """Copy :arr: into shared memory (allocated via :shm_manager: when given, so it is reclaimed on shutdown)."""
        arr = np.ascontiguousarray(arr)
        proto = ArrayProto.from_array(arr)
        nbytes = max(proto.nbytes, 1)
        if shared_memory is None:
            shared_memory = (shm_manager.SharedMemory(size=nbytes) if shm_manager is not None
                             else mp.shared_memory.SharedMemory(create=True, size=nbytes))
        proto.make_from_buffer(shared_memory.buf)[...] = arr
        return cls(proto, shared_memory)

    def __repr__(self):
        return f"{super().__repr__()}; shared_memory={self.shared_memory}"


class SharedArrays:
    """dict-like {key -> SharedArray}; the index lives in a Manager dict, the data in shared-memory segments."""

    def __init__(self, array_headers=None, shm_manager=None):
        assert array_headers is None or isinstance(array_headers, mp.managers.DictProxy)
        assert shm_manager is None or isinstance(shm_manager, mp.managers.SharedMemoryManager)
        if array_headers is None:
            self.array_headers_manager = mp.Manager()
            array_headers = self.array_headers_manager.dict()
        if shm_manager is None:
            shm_manager = mp.managers.SharedMemoryManager()
            shm_manager.start()
        self.array_headers, self.shm_manager = array_headers, shm_manager
        self._attached = {}  # shm name -> SharedMemory, per-process cache

    def fork(self) -> "SharedArrays":
        """A second handle onto the same index and segments."""
        return SharedArrays(self.array_headers, self.shm_manager)

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_attached"] = {}
        state.pop("array_headers_manager", None)
        return state

    def _attach(self, name: str) -> mp.shared_memory.SharedMemory:
        shm = self._attached.get(name)
        if shm is None:
            shm = self._attached[name] = mp.shared_memory.SharedMemory(name=name)
        return shm

    def __getitem__(self, key) -> SharedArray:
        proto, shm_name = self.array_headers[key]
        return SharedArray(proto, self._attach(shm_name))

    def __setitem__(self, key, arr: SharedArray):
        if not isinstance(arr, SharedArray):
            raise ValueError("only SharedArray values can be stored; use create_array(key, proto) and copy into it")
        self._attached.setdefault(arr.shared_memory.name, arr.shared_memory)
        self.array_headers[key] = (ArrayProto.from_array(arr), arr.shared_memory.name)

    def __delitem__(self, key):
        del self.array_headers[key]

    def __contains__(self, key):
        return key in self.array_headers

    def __len__(self):
        return len(self.array_headers)

    def keys(self):
        return self.array_headers.keys()

    def __repr__(self):
        return repr({key: self[key] for key in self.keys()})

    def create_array(self, key, proto: ArrayProto) -> SharedArray:
        """Allocate (through the manager) and register a new array; an existing key is overwritten."""
        arr = SharedArray(proto, self.shm_manager.SharedMemory(size=max(proto.nbytes, 1)))
        self[key] = arr
        return arr
