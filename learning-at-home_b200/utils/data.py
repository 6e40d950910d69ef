"""Tiny tensor helpers (parity: /root/reference/lib/utils/data.py:5-13)."""
import numpy as np
import torch


def check_numpy(x) -> np.ndarray:
    """Return x as a numpy array; tensors are detached and copied to host (this synchronises the device)."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


#: an empty tensor that requires grad: passing it through an autograd.Function forces autograd to call that
#: function's backward even when no real input requires grad (remote experts must always see the backward pass).
DUMMY = torch.empty(0, requires_grad=True)
