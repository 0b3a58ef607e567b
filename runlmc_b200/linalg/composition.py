"""Mirror of runlmc/linalg/composition.py."""
from .matrix import Matrix


class Composition(Matrix):
    """mats[0] mats[1] ... mats[-1], applied right to left (composition.py:14-22)."""

    def __init__(self, mats):
        super().__init__(mats[0].shape[0], mats[-1].shape[1])
        self.mats = mats

    def _apply_dev(self, X):
        for M in reversed(self.mats):
            X = M._apply_dev(X)
        return X
