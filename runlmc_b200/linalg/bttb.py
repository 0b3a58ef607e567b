"""Mirror of runlmc/linalg/bttb.py (symmetric block-Toeplitz with Toeplitz blocks)."""
import numpy as np

from .matrix import Matrix
from .toeplitz import _BTTBHandle


class BTTB(Matrix):
    """:param top: flattened first row, :param sizes: grid sizes n_p (bttb.py:91-108).
    The device path supports up to 3 grid dimensions.
    :raises ValueError: on shape mismatches / empty input (bttb.py:93-101)."""

    def __init__(self, top, sizes):
        top = np.asarray(top)
        sizes = np.asarray(sizes)
        if top.shape != (len(top),):
            raise ValueError('top shape {} is not 1D'.format(top.shape))
        if not top.size:
            raise ValueError('top is empty')
        if sizes.shape != (len(sizes),):
            raise ValueError('sizes shape {} is not 1D'.format(sizes.shape))
        if np.prod(sizes) != top.size:
            raise ValueError("sizes {} don't match grid size {}".format(sizes, top.size))
        if len(sizes) > 3:
            raise ValueError('BTTB on the device supports at most 3 grid dimensions')
        super().__init__(len(top), len(top))
        self.top = top.astype('float64', casting='safe')
        self._sizes = sizes
        self._dev_h = None

    def _handle(self):
        if self._dev_h is None:
            self._dev_h = _BTTBHandle(self.top, [int(s) for s in self._sizes])
        return self._dev_h

    def _apply_dev(self, X):
        return self._handle().apply(X)

    def as_numpy(self):
        sizes = [int(s) for s in self._sizes]
        top = self.top.reshape(sizes)
        idx = np.indices(sizes).reshape(len(sizes), -1)
        diff = np.abs(idx[:, :, None] - idx[:, None, :])
        return top[tuple(diff)]

    def __str__(self):
        if len(self.top) > 50:
            return 'BTTB on grid shape {}'.format(len(self._sizes))
        return 'BTTB on grid \n' + str(self.top.reshape(self._sizes))
