"""
Array / tensor schemas: frozen descriptions of shape+dtype used to pre-allocate buffers and to describe the inputs and
outputs of an expert over the wire (parity: /root/reference/lib/utils/proto.py:6-83).
"""
from dataclasses import dataclass, fields
from typing import Optional, Tuple

import numpy as np
import torch

DUMMY_BATCH_SIZE = 3  # batch size of the throw-away run used to infer an expert's output schema


@dataclass(frozen=True)
class ProtoBase:
    def _fields_dict(self):
        return {f.name: getattr(self, f.name) for f in fields(self)}


@dataclass(frozen=True)
class ArrayProto(ProtoBase):
    shape: tuple
    dtype: np.dtype
    strides: Optional[tuple] = None
    order: str = "C"

    @classmethod
    def from_array(cls, arr: np.ndarray):
        return cls(arr.shape, arr.dtype, strides=arr.strides, order="F" if np.isfortran(arr) else "C")

    def make_empty(self, **overrides) -> np.ndarray:
        return np.ndarray(**{**self._fields_dict(), **overrides})

    def make_from_buffer(self, buffer, offset: int = 0) -> np.ndarray:
        return np.ndarray(self.shape, self.dtype, buffer, offset, strides=self.strides, order=self.order)

    @property
    def nbytes(self) -> int:
        return int(np.dtype(self.dtype).itemsize * np.prod(self.shape))


@dataclass(frozen=True)
class TensorProto(ProtoBase):
    size: tuple
    dtype: torch.dtype = None
    layout: torch.layout = torch.strided
    device: torch.device = None
    requires_grad: bool = False
    pin_memory: bool = False

    @property
    def shape(self):
        return self.size

    @classmethod
    def from_tensor(cls, tensor: torch.Tensor):
        return cls(tuple(tensor.shape), tensor.dtype, tensor.layout, tensor.device, tensor.requires_grad,
                   tensor.is_pinned())

    def make_empty(self, **overrides) -> torch.Tensor:
        return torch.empty(**{**self._fields_dict(), **overrides})

    def convert_array_to_tensor(self, array: np.ndarray) -> torch.Tensor:
        tensor = torch.as_tensor(array, dtype=self.dtype, device=self.device)
        tensor = tensor.requires_grad_(self.requires_grad).to(self.device, non_blocking=True)
        return tensor.pin_memory() if self.pin_memory else tensor


@dataclass(frozen=True, init=False)
class BatchTensorProto(TensorProto):
    """A tensor schema whose 0-th dimension is the (variable) batch size: ``size == (None, *instance_size)``."""

    def __init__(self, *instance_size, **kwargs):
        if len(instance_size) == 1 and isinstance(instance_size[0], (list, tuple, torch.Size)):
            instance_size = tuple(instance_size[0])  # BatchTensorProto((512, 1024)) == BatchTensorProto(512, 1024)
        TensorProto.__init__(self, (None, *instance_size), **kwargs)

    @classmethod
    def from_tensor(cls, tensor: torch.Tensor):
        return cls(*tensor.shape[1:], dtype=tensor.dtype, layout=tensor.layout, device=tensor.device,
                   requires_grad=tensor.requires_grad, pin_memory=tensor.is_pinned())

    def make_empty(self, batch_size: int, **overrides) -> torch.Tensor:
        assert self.size[0] is None, "0-th dimension must be unspecified (None)"
        return super().make_empty(size=(batch_size, *self.size[1:]), **overrides)
