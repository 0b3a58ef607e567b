"""Mirror of runlmc/linalg/toeplitz.py (symmetric Toeplitz via circulant embedding)."""
import ctypes

import numpy as np
import scipy.linalg as la

from .matrix import Matrix
from .. import _native as nat
from .. import device as dev

EPS = np.finfo('float64').eps


class _BTTBHandle:
    """Owns an lmc_bttb (spectrum of the embedded top row on the device)."""

    def __init__(self, top, sizes):
        self.h = ctypes.c_void_p()
        nat.require_cuda()
        sizes = nat.as_i32(sizes)
        top = nat.as_f64(top)
        nat.check(nat.lib.lmc_bttb_create(ctypes.byref(self.h), len(sizes),
                                          nat.host_ptr(sizes), nat.host_ptr(top)))

    def apply(self, X):
        Y = dev.empty(tuple(X.shape))
        nat.check(nat.lib.lmc_bttb_apply(self.h, dev.ptr(X), X.shape[0], dev.ptr(Y), dev.stream()))
        return Y

    def __del__(self):
        h, self.h = getattr(self, 'h', None), None
        if h:
            try:
                nat.lib.lmc_bttb_destroy(h)
            except Exception:
                pass


class Toeplitz(Matrix):
    """Symmetric Toeplitz matrix given by its first row (toeplitz.py:17-44).

    The reference embeds into a circulant of length exactly 2n; here the
    embedding is the next power of two >= 2n with the same structure
    [t, 0.., t[n-1:0:-1]], which yields the same product.

    :raises ValueError: if `top` isn't 1-D or is empty."""

    def __init__(self, top):
        top = np.asarray(top)
        if top.shape != (len(top),):
            raise ValueError('top shape {} is not 1D'.format(top.shape))
        if not top.size:
            raise ValueError('top is empty')
        super().__init__(len(top), len(top))
        self.top = top.astype('float64', casting='safe')
        self._dev_h = None

    def _handle(self):
        if self._dev_h is None:
            self._dev_h = _BTTBHandle(self.top, [len(self.top)])
        return self._dev_h

    def _apply_dev(self, X):
        return self._handle().apply(X)

    def as_numpy(self):
        return la.toeplitz(self.top)

    def upper_eig_bound(self):
        """Gershgorin bound: the largest absolute row sum (what toeplitz.py:69-85 returns).  Row i of a
        symmetric Toeplitz matrix holds t_i .. t_1 t_0 t_1 .. t_{n-1-i}, so its absolute sum is
        P[i] + P[n-1-i] - |t_0| with P the prefix sums of |t|."""
        prefix = np.cumsum(np.abs(self.top))
        rows = prefix + prefix[::-1] - abs(self.top[0])
        return rows.max() * (1 + EPS * len(self.top))

    def __str__(self):
        topstr = 'size {}'.format(len(self.top)) if len(self.top) > 10 else str(self.top)
        return 'Toeplitz ' + topstr
