"""Device-buffer plumbing for the runlmc.linalg mirror: torch owns the memory
and the stream, liblmc_b200.so does the arithmetic."""
import ctypes

import numpy as np

from . import _native as nat


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def to_device(a, dtype=None):
    torch = nat.require_cuda()
    a = np.ascontiguousarray(a, dtype=dtype or np.float64)
    return torch.as_tensor(a, device='cuda')


def empty(shape):
    torch = nat.require_cuda()
    return torch.empty(shape, dtype=torch.float64, device='cuda')


def stream():
    return nat.current_stream_ptr()
