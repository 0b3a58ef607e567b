"""The small composite operators of runlmc.linalg -- Identity, Diag, SumMatrix, Composition,
BlockDiag, SymmSquareBlockMatrix -- on device blocks.

One module because they share their plumbing: each node maps a device block X [k, cols] to
[k, rows] (`_apply_dev`) and the host entry points (matvec / matmat / as_linear_operator) come
from `Matrix`.  Constructor arguments, public attributes and `ValueError`s are the reference's
(identity.py:15-19, diag.py:17-30, sum_matrix.py:19-32, composition.py:14-22,
block_diag.py:24-40, block_matrix.py:12-37); the per-class modules of the reference's layout
re-export from here."""
import numpy as np
import scipy.linalg

from .matrix import Matrix
from .. import _native as nat
from .. import device as dev


def _add_into(total, Y):
    """total += Y on the device (SumMatrix / block-row accumulation)."""
    nat.check(nat.lib.lmc_axpby(Y.numel(), 1.0, dev.ptr(Y), 1.0, dev.ptr(total), dev.stream()))


def _offsets(lens):
    """(starts, stops) of consecutive segments with the given lengths."""
    stops = np.cumsum(np.asarray(lens, dtype=np.int64))
    return stops - np.asarray(lens, dtype=np.int64), stops


def _listing(title, label, items):
    return title + '\n' + '\n'.join('{}{}\n{!s}'.format(label, i, it) for i, it in enumerate(items))


class Identity(Matrix):
    def __init__(self, n):
        super().__init__(n, n)

    # the identity hands its argument back on every level of the protocol
    def matvec(self, x):
        return x

    matmat = matvec
    _apply_dev = matvec

    def as_numpy(self):
        return np.eye(self.shape[0])

    def upper_eig_bound(self):
        return 1


class Diag(Matrix):
    """diag(v).  :raises ValueError: unless v is a vector."""

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
        k, length = X.shape
        Y = dev.empty((k, length))
        nat.check(nat.lib.lmc_diag_apply(dev.ptr(self._dev_v), length, dev.ptr(X), k, dev.ptr(Y), dev.stream()))
        return Y

    def as_numpy(self):
        return np.diag(self.v)

    def upper_eig_bound(self):
        return self.v.max()

    def __str__(self):
        return 'Diag(len {}): {}'.format(len(self.v), self.v)


class SumMatrix(Matrix):
    """K_1 + ... + K_r.  :raises ValueError: for an empty list or differing shapes."""

    def __init__(self, Ks):
        if not Ks:
            raise ValueError('Need at least one matrix to sum')
        shapes = [K.shape for K in Ks]
        if any(s != shapes[0] for s in shapes):
            raise ValueError('At most one distinct shape expected in sum, '
                             'found shapes:\n{}'.format(shapes))
        super().__init__(*shapes[0])
        self.Ks = Ks

    def _apply_dev(self, X):
        total = self.Ks[0]._apply_dev(X)
        if total.data_ptr() == X.data_ptr():
            total = total.clone()          # an Identity term handed its input back: results are fresh blocks
        for K in self.Ks[1:]:
            _add_into(total, K._apply_dev(X))
        return total

    def as_numpy(self):
        return sum(K.as_numpy() for K in self.Ks)

    def upper_eig_bound(self):
        return sum(K.upper_eig_bound() for K in self.Ks)

    def __str__(self):
        return _listing('SumMatrix([..., Ki, ...])', 'K', self.Ks)


class Composition(Matrix):
    """mats[0] @ mats[1] @ ... @ mats[-1]: the last factor meets the vector first."""

    def __init__(self, mats):
        super().__init__(mats[0].shape[0], mats[-1].shape[1])
        self.mats = mats

    def _apply_dev(self, X):
        for factor in self.mats[::-1]:
            X = factor._apply_dev(X)
        return X


class BlockDiag(Matrix):
    """Direct sum of blocks, which may be rectangular."""

    def __init__(self, blocks):
        heights = [b.shape[0] for b in blocks]
        widths = [b.shape[1] for b in blocks]
        super().__init__(sum(heights), sum(widths))
        self.rbegins, self.rends = _offsets(heights)
        self.cbegins, self.cends = _offsets(widths)
        self.blocks = blocks

    def _apply_dev(self, X):
        out = dev.empty((X.shape[0], self.shape[0]))
        for i, blk in enumerate(self.blocks):
            piece = X[:, self.cbegins[i]:self.cends[i]].contiguous()
            out[:, self.rbegins[i]:self.rends[i]] = blk._apply_dev(piece)
        return out

    def as_numpy(self):
        return scipy.linalg.block_diag(*[b.as_numpy() for b in self.blocks])

    def __str__(self):
        return _listing('BlockDiag(..., blocki, ...)', 'block', self.blocks)


class SymmSquareBlockMatrix(Matrix):
    """D x D arrangement of square blocks of one size.  :raises ValueError: on uneven sizes."""

    def __init__(self, blocks):
        self.D = len(blocks)
        if any(len(row) != self.D for row in blocks):
            raise ValueError('Uneven sizes')
        m = blocks[0][0].shape[0]
        super().__init__(self.D * m, self.D * m)
        self.blocks = blocks
        self.begins = m * np.arange(self.D)
        self.ends = self.begins + m

    def _apply_dev(self, X):
        out = dev.empty(tuple(X.shape))
        pieces = [X[:, b:e].contiguous() for b, e in zip(self.begins, self.ends)]
        for i, row in enumerate(self.blocks):
            acc = row[0]._apply_dev(pieces[0])
            if acc.data_ptr() == pieces[0].data_ptr():
                acc = acc.clone()          # never accumulate into a column block that later rows read
            for blk, piece in zip(row[1:], pieces[1:]):
                _add_into(acc, blk._apply_dev(piece))
            out[:, self.begins[i]:self.ends[i]] = acc
        return out

    def as_numpy(self):
        return np.block([[b.as_numpy() for b in row] for row in self.blocks])

    def upper_eig_bound(self):
        bounds = np.array([[b.upper_eig_bound() for b in row] for row in self.blocks], dtype=float)
        return scipy.linalg.norm(bounds, 1)

    def __str__(self):
        return 'SymmBlockMatrix({0} x {0} blocks)'.format(self.D)
