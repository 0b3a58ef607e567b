"""Mirror of runlmc/linalg/block_matrix.py."""
import numpy as np
import scipy.linalg as la

from .matrix import Matrix
from .. import _native as nat
from .. import device as dev


class SymmSquareBlockMatrix(Matrix):
    """D x D array of equally sized square blocks (block_matrix.py:12-37).
    :raises ValueError: on uneven sizes."""

    def __init__(self, blocks):
        self.D = len(blocks)
        if set(map(len, blocks)) != {self.D}:
            raise ValueError('Uneven sizes')
        m = blocks[0][0].shape[0]
        n = self.D * m
        super().__init__(n, n)
        self.blocks = blocks
        self.begins = np.arange(0, self.shape[0], m)
        self.ends = self.begins + m

    def _apply_dev(self, X):
        out = dev.empty(tuple(X.shape))
        cols = [X[:, b:e].contiguous() for b, e in zip(self.begins, self.ends)]
        for rb, re, row in zip(self.begins, self.ends, self.blocks):
            acc = None
            for xc, blk in zip(cols, row):
                Y = blk._apply_dev(xc)
                if acc is None:
                    acc = Y
                else:
                    nat.check(nat.lib.lmc_axpby(Y.numel(), 1.0, dev.ptr(Y), 1.0, dev.ptr(acc), dev.stream()))
            out[:, rb:re] = acc
        return out

    def as_numpy(self):
        z = np.zeros(self.shape)
        for rb, re, row in zip(self.begins, self.ends, self.blocks):
            for cb, ce, blk in zip(self.begins, self.ends, row):
                z[rb:re, cb:ce] = blk.as_numpy()
        return z

    def upper_eig_bound(self):
        bounds = np.array([[b.upper_eig_bound() for b in row] for row in self.blocks], dtype=float)
        return la.norm(bounds, 1)

    def __str__(self):
        return 'SymmBlockMatrix({0} x {0} blocks)'.format(self.D)
