"""Mirror of runlmc/linalg/matrix.py: the lazy linear-operator protocol.

Same public surface (shape, dtype, matvec, matmat, as_numpy,
as_linear_operator, is_square, Matrix.wrap, pickling hooks; reference
matrix.py:7-90).  Every concrete class implements ``_apply_dev``: a block
product on device buffers executed by liblmc_b200.so; ``matvec``/``matmat``
move numpy data to the device, call it, and bring the result back.  There is
no CPU implementation."""
import numpy as np
import scipy.sparse.linalg

from .. import device as dev


class Matrix:
    """:param n: number of rows, :param m: number of columns
    :raises ValueError: if n < 1 or m < 1 (matrix.py:19-21)"""

    def __init__(self, n, m):
        if n < 1 or m < 1:
            raise ValueError('Size of the matrix {} < 1'.format((n, m)))
        self.dtype = np.float64
        self.shape = (n, m)
        self._op = None

    # -- device protocol -------------------------------------------------
    def _apply_dev(self, X):
        """X: float64 CUDA tensor [k, cols] (contiguous) -> [k, rows]."""
        raise NotImplementedError

    # -- reference protocol ----------------------------------------------
    def as_linear_operator(self):
        if self._op is None:
            self._op = scipy.sparse.linalg.LinearOperator(
                shape=self.shape, dtype=self.dtype,
                matvec=self.matvec, matmat=self.matmat)
        return self._op

    def as_numpy(self):
        return self.matmat(np.identity(self.shape[1]))

    def matvec(self, x):
        x = np.asarray(x)
        X = dev.to_device(x.reshape(1, -1))
        if X.shape[1] != self.shape[1]:
            raise ValueError('dimension mismatch: {} vs {}'.format(x.shape, self.shape))
        return self._apply_dev(X).cpu().numpy().reshape(-1)

    def matmat(self, X):
        X = np.asarray(X)
        if X.ndim != 2 or X.shape[0] != self.shape[1]:
            raise ValueError('dimension mismatch: {} vs {}'.format(X.shape, self.shape))
        Xd = dev.to_device(np.ascontiguousarray(X.T))
        return np.ascontiguousarray(self._apply_dev(Xd).cpu().numpy().T)

    def is_square(self):
        return self.shape[0] == self.shape[1]

    @staticmethod
    def wrap(shape, mvm):
        return _MatrixImpl(shape, mvm)

    def __getstate__(self):
        # device handles / cached operators are rebuilt lazily after unpickling
        state = self.__dict__.copy()
        state['_op'] = None
        for k in list(state):
            if k.startswith('_dev_'):
                state[k] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)


class _MatrixImpl(Matrix):
    """Matrix.wrap: a user-supplied host callback (matrix.py:84-90)."""

    def __init__(self, shape, mvm):
        super().__init__(*shape)
        self._mvm = mvm

    def matvec(self, x):
        return self._mvm(x)

    def _apply_dev(self, X):
        host = X.cpu().numpy()
        return dev.to_device(np.array([self._mvm(row) for row in host]))
