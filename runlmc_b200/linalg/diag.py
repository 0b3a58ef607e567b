"""Mirror of runlmc/linalg/diag.py."""
import numpy as np

from .matrix import Matrix
from .. import _native as nat
from .. import device as dev


class Diag(Matrix):
    """:raises ValueError: if v is not a non-empty vector (diag.py:17-22)."""

    def __init__(self, v):
        v = np.asarray(v)
        super().__init__(len(v), len(v))
        if v.ndim != 1:
            raise ValueError('Expected input vector for Diagonal matrix '
                             'go something of shape {}'.format(v))
        self.v = v
        self._dev_v = None

    def _apply_dev(self, X):
        if self._dev_v is None:
            self._dev_v = dev.to_device(self.v)
        Y = dev.empty(tuple(X.shape))
        nat.check(nat.lib.lmc_diag_apply(dev.ptr(self._dev_v), X.shape[1], dev.ptr(X), X.shape[0],
                                         dev.ptr(Y), dev.stream()))
        return Y

    def as_numpy(self):
        return np.diag(self.v)

    def __str__(self):
        return 'Diag(len {}): {}'.format(len(self.v), self.v)

    def upper_eig_bound(self):
        return self.v.max()
