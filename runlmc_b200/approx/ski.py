"""Mirror of runlmc/approx/ski.py: W K W^T as a composition."""
import numpy as np

from ..linalg.composition import Composition
from ..linalg.matrix import Matrix
from .. import _native as nat
from .. import device as dev


class _DeviceCSR(Matrix):
    """scipy CSR product on the device (replaces Matrix.wrap(W.shape, W.dot), ski.py:12-17)."""

    def __init__(self, csr):
        super().__init__(*csr.shape)
        self.csr = csr.tocsr()
        self._dev_arrays = None

    def _arrays(self):
        if self._dev_arrays is None:
            c = self.csr
            self._dev_arrays = (dev.to_device(c.indptr, np.int32), dev.to_device(c.indices, np.int32),
                                dev.to_device(c.data, np.float64))
        return self._dev_arrays

    def _apply_dev(self, X):
        indptr, indices, data = self._arrays()
        k = X.shape[0]
        Y = dev.empty((k, self.shape[0]))
        nat.check(nat.lib.lmc_csr_apply(self.shape[0], dev.ptr(indptr), dev.ptr(indices), dev.ptr(data),
                                        dev.ptr(X), X.shape[1], k, dev.ptr(Y), self.shape[0], dev.stream()))
        return Y

    def as_numpy(self):
        return self.csr.toarray()


class SKI(Composition):
    """:param K: grid kernel (a Matrix), :param W: interpolant CSR, :param WT: its transpose CSR."""

    def __init__(self, K, W, WT):
        self.W = W
        self.K = K
        self.WT = WT
        super().__init__([_DeviceCSR(W), K, _DeviceCSR(WT)])

    def as_numpy(self):
        WKT = self.W.dot(self.K.as_numpy().T)
        return self.W.dot(WKT.T)

    def upper_eig_bound(self):
        return self.K.upper_eig_bound() * self.shape[0] / self.K.shape[0]
