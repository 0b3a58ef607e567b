"""Mirror of runlmc/approx/interpolation.py: cubic-convolution interpolants.

Same functions and return types (scipy CSR) as the reference, so callers and
tests are unchanged.  These CSR matrices are host-side descriptions; the device
operator does not read them: `multi_interpolant` attaches the geometry
(points + grids) to the CSR it returns and the fused CUDA operator
(runlmc_b200.fused.FusedLMC) recomputes the weights from coordinates.
"""
import logging

import numpy as np
import scipy.sparse

_LOG = logging.getLogger(__name__)


def cubic_kernel(x):
    """Keys' cubic convolution kernel, supported on |x| <= 2
    (interpolation.py:21-53).  :raises ValueError: if any |x| > 2."""
    x = np.fabs(np.asarray(x, dtype=float))
    if np.any(x > 2):
        raise ValueError('only absolute values <= 2 allowed')
    inner = ((1.5 * x - 2.5) * x) * x + 1
    outer = ((-0.5 * x + 2.5) * x - 4) * x + 2
    return np.where(x <= 1, inner, outer)


def _axis_stencil(grid, s):
    """Clamped columns [n,4] and weights [n,4] of the 4-tap stencil along one axis
    (interpolation.py:98-115): taps c in (-2,-1,0,1) -> column i0-c, weight k(u+c)."""
    m = len(grid)
    delta = grid[1] - grid[0]
    f = (s - grid[0]) / delta
    i0 = np.floor(f)
    u = f - i0
    taps = np.array([-2, -1, 0, 1])
    cols = np.clip(i0[:, None] - taps[None, :], 0, m - 1).astype(np.int64)
    w = cubic_kernel(u[:, None] + taps[None, :])
    return cols, w


def _check_grid(name, grid):
    if grid.ndim != 1:
        raise ValueError('{} dim {} should be 1'.format(name, grid.ndim))
    if grid.size < 4:
        raise ValueError('grid size {} must be >=4'.format(grid.size))


def _warn_range(axis, s, grid):
    if s.min() <= grid[0] or s.max() >= grid[-1]:
        _LOG.warning('%srange of samples [%f, %f] outside grid range [%f, %f]',
                     axis, s.min(), s.max(), grid[0], grid[-1])


def interp_cubic(grid, samples):
    """n x m CSR interpolation matrix, <= 4 entries per row (interpolation.py:56-116)."""
    grid = np.asarray(grid)
    samples = np.asarray(samples)
    m = len(grid)
    n = samples.size
    if n == 0:
        return scipy.sparse.csr_matrix((0, m), dtype=float)
    if grid.ndim != 1:
        raise ValueError('grid dim {} should be 1'.format(grid.ndim))
    if samples.ndim != 1:
        raise ValueError('samples dim {} should be 1'.format(samples.ndim))
    if m < 4:
        raise ValueError('grid size {} must be >=4'.format(m))
    _warn_range('', samples, grid)
    cols, w = _axis_stencil(grid, samples)
    rows = np.repeat(np.arange(n), 4)
    out = scipy.sparse.coo_matrix((w.ravel(), (rows, cols.ravel())), shape=(n, m)).tocsr()
    out.sum_duplicates()
    return out


def interp_bicubic(gridx, gridy, samples):
    """n x (mx*my) CSR, <= 16 entries per row, column ix*my + iy
    (interpolation.py:218-328)."""
    gridx = np.asarray(gridx)
    gridy = np.asarray(gridy)
    samples = np.asarray(samples)
    mx, my = gridx.size, gridy.size
    n = samples.shape[0]
    if n == 0:
        return scipy.sparse.csr_matrix((0, mx * my), dtype=float)
    _check_grid('gridx', gridx)
    _check_grid('gridy', gridy)
    if samples.ndim != 2 or samples.shape[1] != 2:
        raise ValueError('expecting 2d samples, got shape {}'.format(samples.shape))
    _warn_range('x ', samples[:, 0], gridx)
    _warn_range('y ', samples[:, 1], gridy)
    cx, wx = _axis_stencil(gridx, samples[:, 0])
    cy, wy = _axis_stencil(gridy, samples[:, 1])
    cols = (cx[:, :, None] * my + cy[:, None, :]).reshape(n, 16)
    w = (wy[:, None, :] * wx[:, :, None]).reshape(n, 16)
    rows = np.repeat(np.arange(n), 16)
    out = scipy.sparse.coo_matrix((w.ravel(), (rows, cols.ravel())), shape=(n, mx * my)).tocsr()
    out.sum_duplicates()
    return out


class LazyCSR(scipy.sparse.csr_matrix):
    """A scipy CSR matrix whose data / indices / indptr are built by `build()` the first time anything
    reads them.  The fused device operator recomputes the interpolation weights from the coordinates
    and never reads the CSR, so the reference's call sequence
    ``W = multi_interpolant(Xs, *grids); WT = W.transpose().tocsr()`` (interpolated_llgp.py:431-437)
    costs nothing on the host until a caller actually looks inside W (at n = 1M the eager build takes
    seconds and 16 M nonzeros)."""

    def __init__(self, arg1, shape=None, dtype=None, copy=False, build=None):
        self.__dict__['_lazy'] = None
        if build is None:       # scipy's own constructor calls (self.__class__(other) inside binary operators)
            scipy.sparse.csr_matrix.__init__(self, arg1, shape=shape, dtype=dtype, copy=copy)
            return
        scipy.sparse.csr_matrix.__init__(self, (int(arg1[0]), int(arg1[1])), dtype=np.float64)
        self.__dict__['_lazy'] = build      # the base constructor stored empty arrays: replaced on first read

    def _materialize(self):
        build = self.__dict__.get('_lazy')
        if build is not None:
            self.__dict__['_lazy'] = None
            real = build()
            self.__dict__['_d'], self.__dict__['_i'], self.__dict__['_p'] = real.data, real.indices, real.indptr

    @property
    def materialized(self):
        return self.__dict__.get('_lazy') is None

    def _get(self, key):
        self._materialize()
        return self.__dict__[key]

    data = property(lambda self: self._get('_d'), lambda self, v: self.__dict__.__setitem__('_d', v))
    indices = property(lambda self: self._get('_i'), lambda self, v: self.__dict__.__setitem__('_i', v))
    indptr = property(lambda self: self._get('_p'), lambda self, v: self.__dict__.__setitem__('_p', v))

    def transpose(self, axes=None, copy=False):
        if self.materialized:
            return scipy.sparse.csr_matrix((self.data, self.indices, self.indptr), shape=self.shape).transpose(axes, copy)
        return _LazyTranspose(self)

    def __reduce__(self):
        # pickles as the plain CSR it stands for (plus its attributes, e.g. lmc_geometry)
        plain = scipy.sparse.csr_matrix((self.data, self.indices, self.indptr), shape=self.shape)
        extra = {k: v for k, v in self.__dict__.items() if k in ('lmc_geometry',)}
        return (_restore_csr, (plain, extra))


def _restore_csr(plain, extra):
    plain.__dict__.update(extra)
    return plain


class _LazyTranspose:
    """`W.transpose()` of a LazyCSR that has not been built: only `.tocsr()` / `.shape` are offered,
    which is what the reference's call sequence uses."""

    def __init__(self, W):
        self._W = W
        self.shape = (W.shape[1], W.shape[0])

    def tocsr(self, copy=False):
        W = self._W

        def build():
            return scipy.sparse.csr_matrix((W.data, W.indices, W.indptr), shape=W.shape).transpose().tocsr()

        return LazyCSR(self.shape, build=build)

    def __getattr__(self, name):     # anything else: fall back to the real transposed matrix
        W = self._W
        real = scipy.sparse.csr_matrix((W.data, W.indices, W.indptr), shape=W.shape).transpose()
        return getattr(real, name)


def multi_interpolant(Xs, *inducing_grids):
    """Block-diagonal interpolant over outputs: row block d <-> column block
    [d*m, (d+1)*m) (interpolation.py:119-176).  The result carries
    ``lmc_geometry = (Xs, grids)`` for the fused device operator, and is a LazyCSR: argument checks and
    range warnings happen here, the CSR arrays are assembled only if something reads them."""
    grids = [np.asarray(g) for g in inducing_grids]
    Xs = [np.asarray(X) for X in Xs]
    one_d = Xs[0].ndim == 1 or Xs[0].shape[1] == 1
    m = int(np.prod([g.size for g in grids[:1 if one_d else 2]]))
    n = int(sum(len(X) for X in Xs))
    # the per-output argument checks and warnings of interp_cubic / interp_bicubic, without the assembly
    for g in grids[:1 if one_d else 2]:
        _check_grid('grid', g)
    for X in Xs:
        if len(X) == 0:
            continue
        if one_d:
            _warn_range('', X.ravel(), grids[0])
        else:
            if X.ndim != 2 or X.shape[1] != 2:
                raise ValueError('expecting 2d samples, got shape {}'.format(X.shape))
            _warn_range('x ', X[:, 0], grids[0])
            _warn_range('y ', X[:, 1], grids[1])

    def build():
        with _quiet():
            if one_d:
                Ws = [interp_cubic(grids[0], X.ravel()) for X in Xs]
            else:
                Ws = [interp_bicubic(grids[0], grids[1], X) for X in Xs]
        W = scipy.sparse.block_diag(Ws, format='csr') if len(Ws) > 1 else Ws[0].tocsr()
        return scipy.sparse.csr_matrix(W)

    W = LazyCSR((n, len(Xs) * m), build=build)
    W.lmc_geometry = (Xs, grids)
    return W


class _quiet:
    """The range warnings were already issued by multi_interpolant itself."""

    def __enter__(self):
        self._level = _LOG.level
        _LOG.setLevel(logging.ERROR)

    def __exit__(self, *exc):
        _LOG.setLevel(self._level)


def autogrid(Xs, lo, hi, m):
    """Equispaced grid per dimension covering the data with two extra cells on
    each side; note the result has m + 4 points (interpolation.py:179-215)."""
    P = Xs[0].shape[1]
    assert lo is None or len(lo) == P, (P, lo)
    assert hi is None or len(hi) == P, (P, hi)
    assert m is None or len(m) == P, (P, m)
    data_lo = np.vstack([X.min(axis=0) for X in Xs]).min(axis=0)
    data_hi = np.vstack([X.max(axis=0) for X in Xs]).max(axis=0)
    if m is None:
        m = np.ones(P) * (sum(len(X) for X in Xs) // len(Xs))
    else:
        m = np.array(m, dtype=float)
    lo = (data_lo if lo is None else np.minimum(lo, data_lo)).astype(float)
    hi = (data_hi if hi is None else np.maximum(hi, data_hi)).astype(float)
    delta = (hi - lo) / m
    lo = lo - 2 * delta
    hi = hi + 2 * delta
    return [np.linspace(a, b, int(k) + 4) for a, b, k in zip(lo, hi, m)]
