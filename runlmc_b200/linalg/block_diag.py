"""Mirror of runlmc/linalg/block_diag.py."""
import numpy as np
import scipy.linalg as la

from .matrix import Matrix
from .. import device as dev


def begin_end_indices(lens):
    ends = np.add.accumulate(lens)
    begins = np.roll(ends, 1)
    begins[0] = 0
    return begins, ends


class BlockDiag(Matrix):
    """Direct sum of (possibly rectangular) blocks (block_diag.py:24-40)."""

    def __init__(self, blocks):
        row_lens = [b.shape[0] for b in blocks]
        col_lens = [b.shape[1] for b in blocks]
        super().__init__(sum(row_lens), sum(col_lens))
        self.rbegins, self.rends = begin_end_indices(row_lens)
        self.cbegins, self.cends = begin_end_indices(col_lens)
        self.blocks = blocks

    def _apply_dev(self, X):
        out = dev.empty((X.shape[0], self.shape[0]))
        for rb, re, cb, ce, blk in zip(self.rbegins, self.rends, self.cbegins, self.cends, self.blocks):
            out[:, rb:re] = blk._apply_dev(X[:, cb:ce].contiguous())
        return out

    def as_numpy(self):
        return la.block_diag(*(b.as_numpy() for b in self.blocks))

    def __str__(self):
        return ('BlockDiag(..., blocki, ...)\n' +
                '\n'.join(['block{}\n{!s}'.format(i, b) for i, b in enumerate(self.blocks)]))
