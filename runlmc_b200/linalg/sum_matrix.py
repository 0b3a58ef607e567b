"""Mirror of runlmc/linalg/sum_matrix.py."""
from .matrix import Matrix
from .. import _native as nat
from .. import device as dev


class SumMatrix(Matrix):
    """sum_i K_i.  :raises ValueError: if `Ks` is empty or shapes differ
    (sum_matrix.py:19-26)."""

    def __init__(self, Ks):
        if not Ks:
            raise ValueError('Need at least one matrix to sum')
        shapes = [K.shape for K in Ks]
        if len(set(shapes)) != 1:
            raise ValueError('At most one distinct shape expected in sum, '
                             'found shapes:\n{}'.format(shapes))
        super().__init__(*shapes[0])
        self.Ks = Ks

    def _apply_dev(self, X):
        total = None
        for K in self.Ks:
            Y = K._apply_dev(X)
            if total is None:
                total = Y if Y.data_ptr() != X.data_ptr() else Y.clone()
            else:
                nat.check(nat.lib.lmc_axpby(Y.numel(), 1.0, dev.ptr(Y), 1.0, dev.ptr(total), dev.stream()))
        return total

    def as_numpy(self):
        return sum(K.as_numpy() for K in self.Ks)

    def __str__(self):
        return ('SumMatrix([..., Ki, ...])\n' +
                '\n'.join(['K{}\n{!s}'.format(i, K) for i, K in enumerate(self.Ks)]))

    def upper_eig_bound(self):
        return sum(K.upper_eig_bound() for K in self.Ks)
