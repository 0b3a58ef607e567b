"""Mirror of runlmc/linalg/kronecker.py."""
import numpy as np

from .matrix import Matrix
from .numpy_matrix import NumpyMatrix
from .. import _native as nat
from .. import device as dev


def _transpose(X, k, a, b):
    """[k, a, b] -> [k, b, a] on the device."""
    Y = dev.empty((k, a * b))
    nat.check(nat.lib.lmc_transpose(dev.ptr(X), k, a, b, dev.ptr(Y), dev.stream()))
    return Y


class Kronecker(Matrix):
    """A (x) B: (A (x) B) x = vec(A X B^T), X = x.reshape(cols_A, cols_B)
    (kronecker.py:39-46)."""

    def __init__(self, A, B):
        super().__init__(A.shape[0] * B.shape[0], A.shape[1] * B.shape[1])
        self.A = A
        self.B = B

    def as_numpy(self):
        return np.kron(self.A.as_numpy(), self.B.as_numpy())

    def _apply_dev(self, X):
        k = X.shape[0]
        ra, ca = self.A.shape
        rb, cb = self.B.shape
        # B along the fast axis: [k*ca, cb] -> [k*ca, rb]
        T = self.B._apply_dev(X.reshape(k * ca, cb)).reshape(k, ca * rb)
        if isinstance(self.A, NumpyMatrix):
            return self.A._apply_dev(T, inner=rb)          # contracts the slow axis in place
        # general A: transpose, apply along the (now fast) axis, transpose back
        Tt = _transpose(T, k, ca, rb)                       # [k, rb, ca]
        U = self.A._apply_dev(Tt.reshape(k * rb, ca)).reshape(k, rb * ra)
        return _transpose(U, k, rb, ra)                     # [k, ra, rb]

    def __str__(self):
        return 'Kron(A, B)\nA\n{!s}\nB\n{!s}'.format(self.A, self.B)

    def upper_eig_bound(self):
        return self.A.upper_eig_bound() * self.B.upper_eig_bound()
