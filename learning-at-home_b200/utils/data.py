"""Tiny tensor helpers (parity: /root/reference/lib/utils/data.py:5-13)."""
from typing import Any

import numpy as np
import torch

# numpy has no bfloat16 / float8: tensors of those dtypes are widened before they leave torch
_WIDEN = {torch.bfloat16: torch.float32}
for _name in ("float8_e4m3fn", "float8_e5m2"):
    if hasattr(torch, _name):
        _WIDEN[getattr(torch, _name)] = torch.float32


def check_numpy(x: Any) -> np.ndarray:
    """
    Host numpy view of ``x``.  Tensors are detached and brought to the host first — for a CUDA tensor this is a blocking
    device-to-host copy (it synchronises the launching stream); dtypes numpy cannot represent (bf16, fp8) come back as fp32.
    Anything else goes through ``np.asarray``.
    """
    if not torch.is_tensor(x):
        return np.asarray(x)
    t = x.detach()
    widened = _WIDEN.get(t.dtype)
    if widened is not None:
        t = t.to(widened)
    return t.cpu().numpy()


def _make_dummy() -> torch.Tensor:
    t = torch.zeros(0)
    t.requires_grad_(True)
    return t


#: an empty tensor that requires grad: passing it through an autograd.Function forces autograd to call that
#: function's backward even when no real input requires grad (remote experts must always see the backward pass).
DUMMY = _make_dummy()
