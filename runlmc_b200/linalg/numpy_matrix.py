"""Mirror of runlmc/linalg/numpy_matrix.py (dense adapter)."""
import numpy as np

from .matrix import Matrix
from .. import _native as nat
from .. import device as dev


class NumpyMatrix(Matrix):
    """:raises ValueError: if `nparr` isn't 2-D (numpy_matrix.py:18-23)."""

    def __init__(self, nparr):
        nparr = np.asarray(nparr)
        if nparr.ndim != 2:
            raise ValueError('Input numpy array of shape {} not matrix'.format(nparr.shape))
        self.A = nparr.astype('float64', casting='safe')
        super().__init__(*self.A.shape)
        self._dev_A = None

    def _A(self):
        if self._dev_A is None:
            self._dev_A = dev.to_device(self.A)
        return self._dev_A

    def _apply_dev(self, X, inner=1):
        """Y[k][r][i] = sum_c A[r][c] X[k][c][i]; X is [k, cols*inner]."""
        k = X.shape[0]
        Y = dev.empty((k, self.shape[0] * inner))
        nat.check(nat.lib.lmc_dense_apply(dev.ptr(self._A()), self.shape[0], self.shape[1],
                                          dev.ptr(X), k, inner, dev.ptr(Y), dev.stream()))
        return Y

    def as_numpy(self):
        return self.A

    def __str__(self):
        return str(self.A)

    def upper_eig_bound(self):
        return np.abs(self.A).sum(axis=1).max()
