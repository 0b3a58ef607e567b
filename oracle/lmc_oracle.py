"""CPU oracle for the SKI-LMC MVM + probe-MINRES gradient path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  Nothing under ``runlmc_b200/`` imports it and
the product path never falls back to it.

It is a functional numpy/scipy restatement (not a copy) of the arithmetic that
vlad17/runlmc performs on this path.  Each function cites the reference
``file:line`` it follows (paths relative to the reference checkout).  The
reference is pure Python on top of numpy.fft / scipy.sparse / scipy MINRES, so
this oracle uses the same third-party primitives (numpy pocketfft, scipy CSR
``.dot``) in the same order, which makes it agree with the reference to the
last few ulps.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference modules from ``/root/reference`` in the build container, runs them on
seeded inputs and stores inputs+outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this oracle against those files and
against the known-answer cases restated from the reference's own unit tests
(test_toeplitz.py, test_bttb.py, test_kronecker.py, test_sum_matrix.py,
test_interpolation.py).
"""
import math

import numpy as np
import scipy.sparse
import scipy.sparse.linalg

EPS = np.finfo(np.float64).eps

# ---------------------------------------------------------------------------
# Interpolation (runlmc/approx/interpolation.py)
# ---------------------------------------------------------------------------


def cubic_kernel(x):
    """Keys cubic-convolution kernel; interpolation.py:21-53."""
    x = np.asarray(x, dtype=float)
    a = np.fabs(x)
    if np.any(a > 2):
        raise ValueError('only absolute values <= 2 allowed')
    out = np.zeros_like(a)
    near = a <= 1
    an = a[near]
    out[near] = ((1.5 * an - 2.5) * an) * an + 1
    af = a[~near]
    out[~near] = ((-0.5 * af + 2.5) * af - 4) * af + 2
    return out


def _stencil_1d(grid, s):
    """Base index and fractional offset; interpolation.py:98-101.

    Returns (i0, u): floor((s-g0)/delta) as float array and the remainder."""
    delta = grid[1] - grid[0]
    f = (s - grid[0]) / delta
    i0 = np.floor(f)
    return i0, f - i0


def interp_cubic_csr(grid, samples):
    """n x m CSR with <= 4 nnz per row; interpolation.py:56-116.

    Taps c in {-2,-1,0,1}: column clamp(i0 - c), weight k(u + c); clamped
    duplicates accumulate through CSR addition (interpolation.py:105-115)."""
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
    i0, u = _stencil_1d(grid, samples)
    rows = np.arange(n + 1)
    acc = scipy.sparse.csr_matrix((n, m), dtype=float)
    for c in (-2, -1, 0, 1):
        col = np.clip(i0 - c, 0, m - 1)
        acc = acc + scipy.sparse.csr_matrix(
            (cubic_kernel(u + c), col, rows), shape=(n, m))
    return acc


def interp_bicubic_csr(gridx, gridy, samples):
    """n x (mx*my) CSR, <= 16 nnz/row, column ix*my+iy; interpolation.py:218-328.

    Built, as in the reference, as (n x 4n y-weight matrix) . (4n x m stack of
    x-stencil matrices, one per y tap) so that the stored weight is wy*wx with
    clamped duplicates summed."""
    gridx = np.asarray(gridx)
    gridy = np.asarray(gridy)
    samples = np.asarray(samples)
    mx, my = gridx.size, gridy.size
    n = samples.shape[0]
    if n == 0:
        return scipy.sparse.csr_matrix((0, mx * my), dtype=float)
    for name, g in (('gridx', gridx), ('gridy', gridy)):
        if g.ndim != 1:
            raise ValueError('{} dim {} should be 1'.format(name, g.ndim))
        if g.size < 4:
            raise ValueError('grid size {} must be >=4'.format(g.size))
    if samples.ndim != 2 or samples.shape[1] != 2:
        raise ValueError('expecting 2d samples, got shape {}'.format(
            samples.shape))
    i0y, uy = _stencil_1d(gridy, samples[:, 1])
    i0x, ux = _stencil_1d(gridx, samples[:, 0])
    rows = np.arange(n + 1)
    ymat = scipy.sparse.csr_matrix((n, 4 * n), dtype=float)
    xmats = []
    for cy in (-2, -1, 0, 1):
        iy = np.clip(i0y - cy, 0, my - 1)
        xm = scipy.sparse.csr_matrix((n, mx * my), dtype=float)
        for cx in (-2, -1, 0, 1):
            ix = np.clip(i0x - cx, 0, mx - 1)
            xm = xm + scipy.sparse.csr_matrix(
                (cubic_kernel(ux + cx), ix * my + iy, rows),
                shape=(n, mx * my))
        xmats.append(xm)
        ymat = ymat + scipy.sparse.csr_matrix(
            (cubic_kernel(uy + cy), np.arange(n) + n * (cy + 2), rows),
            shape=(n, 4 * n))
    return ymat.dot(scipy.sparse.vstack(xmats, format='csc'))


def multi_interpolant_csr(Xs, *grids):
    """Block-diagonal stack over outputs; interpolation.py:119-176.

    Row block d = points of output d (input order), column block
    [d*m, (d+1)*m)."""
    m = int(np.prod([len(g) for g in grids]))
    if Xs[0].ndim == 1 or Xs[0].shape[1] == 1:
        Ws = [interp_cubic_csr(grids[0], np.asarray(X).ravel()) for X in Xs]
    else:
        Ws = [interp_bicubic_csr(grids[0], grids[1], X) for X in Xs]
    n = sum(W.shape[0] for W in Ws)
    indptr = [np.zeros(1, dtype=np.int64)]
    indices, data = [], []
    off = 0
    for d, W in enumerate(Ws):
        indptr.append(W.indptr[1:].astype(np.int64) + off)
        indices.append(W.indices.astype(np.int64) + d * m)
        data.append(W.data)
        off += W.indptr[-1]
    return scipy.sparse.csr_matrix(
        (np.concatenate(data), np.concatenate(indices),
         np.concatenate(indptr)), shape=(n, len(Xs) * m))


def autogrid(Xs, lo, hi, m):
    """Per-dim linspace covering the data +-2 cells, m+4 points;
    interpolation.py:179-215."""
    P = Xs[0].shape[1]
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
    m = m + 4
    return [np.linspace(a, b, int(k)) for a, b, k in zip(lo, hi, m)]


# ---------------------------------------------------------------------------
# Structured matrices (runlmc/linalg)
# ---------------------------------------------------------------------------

def toeplitz_spectrum(top):
    """rfft of the length-2n circulant [t, 0, t[n-1:0:-1]]; toeplitz.py:43-52."""
    top = np.asarray(top)
    n = len(top)
    c = np.zeros(2 * n)
    c[:n] = top
    c[n + 1:] = top[1:][::-1]
    return np.fft.rfft(c)


def toeplitz_matvec(spec, x):
    """toeplitz.py:57-67."""
    n = len(x)
    xf = np.fft.rfft(x, n=2 * n)
    xf *= spec
    return np.fft.irfft(xf)[:n]


def pow2_sizes(sizes):
    """bttb.py:16-19 applied to 2*sizes."""
    return [2 ** (int(2 * s - 1).bit_length()) for s in sizes]


def bttb_embed(top, sizes):
    """Symmetric circulant embedding padded to pow2 per dim; bttb.py:110-121."""
    sizes = [int(s) for s in sizes]
    emb = pow2_sizes(sizes)
    ext = np.zeros(emb)
    ext[tuple(slice(0, s) for s in sizes)] = np.asarray(
        top, dtype=float).reshape(sizes)
    # mirror each axis, last axis first, exactly as the reference does
    for ax in range(len(sizes) - 1, -1, -1):
        s, e = sizes[ax], emb[ax]
        dst = [slice(None)] * len(sizes)
        src = [slice(None)] * len(sizes)
        dst[ax] = slice(e - s + 1, e)
        src[ax] = slice(s - 1, 0, -1)
        ext[tuple(dst)] = ext[tuple(src)]
    return ext


def bttb_spectrum(top, sizes):
    """bttb.py:106-108."""
    return np.fft.rfftn(bttb_embed(top, sizes))


def bttb_matvec(spec, sizes, x):
    """bttb.py:144-148."""
    sizes = [int(s) for s in sizes]
    xf = np.fft.rfftn(np.asarray(x, dtype=float).reshape(sizes),
                      s=pow2_sizes(sizes), axes=list(range(len(sizes))))
    xf *= spec
    full = np.fft.irfftn(xf, s=pow2_sizes(sizes), axes=list(range(len(sizes))))
    return full[tuple(slice(0, s) for s in sizes)].ravel()


def bttb_dense(top, sizes):
    """Dense BTTB from its top row, independent of the FFT path (used by the
    known-answer tests; mirrors the meaning of bttb.py:123-142)."""
    sizes = [int(s) for s in sizes]
    top = np.asarray(top, dtype=float).reshape(sizes)
    idx = np.indices(sizes).reshape(len(sizes), -1)          # [P, m]
    diff = np.abs(idx[:, :, None] - idx[:, None, :])          # [P, m, m]
    return top[tuple(diff)]


def matmat_by_columns(matvec, X, rows):
    """Default Matrix.matmat: python loop of matvec; matrix.py:55-67."""
    out = np.empty((X.shape[1], rows))
    for i, col in enumerate(X.T):
        out[i] = matvec(col)
    return out.T


def kron_matvec(A, b_matmat, b_shape, x):
    """(A kron B) x with dense A and an operator B; kronecker.py:39-46."""
    x = x.reshape(-1, b_shape[1]).T
    x = b_matmat(x)                       # [rows_B, D]
    x = x.reshape(-1, A.shape[1]).T
    x = A.dot(x)
    return x.reshape(-1)


# ---------------------------------------------------------------------------
# Stationary kernels on distances (runlmc/kern)
# ---------------------------------------------------------------------------

def kern_eval(kind, params, r):
    """k(r): rbf.py:39-40, matern32.py:39-41, std_periodic.py:44-48."""
    r = np.asarray(r, dtype=float)
    if kind == 'rbf':
        return np.exp(-0.5 * np.square(r) * params[0])
    if kind == 'matern32':
        s = r * np.sqrt(3) * params[0]
        return (1 + s) * np.exp(-s)
    if kind == 'periodic':
        sn = np.sin((np.pi / params[1]) * r)
        return np.exp(-0.5 * np.square(sn) * params[0])
    raise ValueError(kind)


def kern_grad(kind, params, r):
    """[dk/dtheta ...]: rbf.py:50-54, matern32.py:51-57, std_periodic.py:58-67."""
    r = np.asarray(r, dtype=float)
    if kind == 'rbf':
        sq = np.square(r)
        return [np.exp(-0.5 * sq * params[0]) * -0.5 * sq]
    if kind == 'matern32':
        s = r * np.sqrt(3) * params[0]
        ds = r * np.sqrt(3)
        e = np.exp(-s)
        return [(1 + s) * (e * -ds) + ds * e]
    if kind == 'periodic':
        sc = np.pi / params[1] * r
        sn = np.sin(sc)
        dsn = np.cos(sc) * sc
        dsn = dsn * (-1 / params[1] * params[0])
        sq = np.square(sn)
        e = np.exp(-0.5 * sq * params[0])
        return [e * -0.5 * sq, e * -1 * sn * dsn]
    raise ValueError(kind)


class KernelSpec:
    """Minimal stand-in for the paramz-based FunctionalKernel (which needs
    paramz, absent here).  Holds exactly the numbers the hot path consumes;
    single active-dimension group.  functional_kernel.py:225-300."""

    def __init__(self, kinds, kparams, coreg_vecs, coreg_diags, noise):
        self.kinds = list(kinds)
        self.kparams = [np.atleast_1d(np.asarray(p, dtype=float))
                        for p in kparams]
        self.coreg_vecs = [np.atleast_2d(np.asarray(a, dtype=float))
                           for a in coreg_vecs]
        self.coreg_diags = [np.asarray(k, dtype=float) for k in coreg_diags]
        self.noise = np.asarray(noise, dtype=float)
        self.Q = len(self.kinds)
        self.D = len(self.noise)

    def coreg_mats(self):
        """B_q = A_q^T A_q + diag(kappa_q); functional_kernel.py:280-287."""
        return [a.T.dot(a) + np.diag(k)
                for a, k in zip(self.coreg_vecs, self.coreg_diags)]

    def tops(self, dists):
        return [kern_eval(k, p, dists)
                for k, p in zip(self.kinds, self.kparams)]

    def top_grads(self, dists):
        return [kern_grad(k, p, dists)
                for k, p in zip(self.kinds, self.kparams)]


# ---------------------------------------------------------------------------
# SKI-LMC operator (runlmc/lmc/grid_kernel.py, runlmc/approx/ski.py)
# ---------------------------------------------------------------------------

class LmcOperator:
    """K~ = W (sum_q B_q kron T_q) W^T + diag(noise repeated).

    The grid part is evaluated in one of the reference's three equivalent
    representations (grid_kernel.py:22-41): 'sum' (:126-136), 'bt' (:115-123)
    or 'slfm' (:77-112); ``rep='auto'`` applies the reference's selection rule
    (grid_kernel.py:52-64).  Noise Diag appended as grid_kernel.py:70-74; SKI
    composition ski.py:8-17; sums evaluated like sum_matrix.py:31-32 (python
    ``sum`` starting from 0)."""

    def __init__(self, W, WT, Bs, tops, sizes, noise, lens, coreg_vecs=None,
                 coreg_diags=None, rep='sum', slfm_identity=(False, False)):
        self.W, self.WT = W, WT
        self.Bs = [np.asarray(B, dtype=float) for B in Bs]
        self.sizes = [int(s) for s in sizes]
        self.m = int(np.prod(self.sizes))
        self.tops = [np.asarray(t, dtype=float).ravel() for t in tops]
        self.specs = [bttb_spectrum(t, self.sizes) for t in self.tops]
        self.noise_rep = np.repeat(np.asarray(noise, dtype=float), lens)
        self.lens = list(lens)
        self.n = W.shape[0]
        self.shape = (self.n, self.n)
        self.D = len(lens)
        Q = len(self.Bs)
        if rep == 'auto':
            if Q == 1:
                rep = 'sum'
            else:
                tot_rank = sum(len(a) for a in coreg_vecs)
                rep = 'slfm' if tot_rank + self.D < self.D ** 2 else 'bt'
        self.rep = rep
        # grid_kernel.py:84-86 / 101-103: with no non-independent kernel the coreg part of the slfm
        # representation is Identity(D m); with neither LMC nor independent kernels the diag part is
        self._coreg_identity, self._diag_identity = slfm_identity
        D = self.D
        if rep == 'bt':
            # grid_kernel.py:115-123: one BTTB per (i <= j) block
            bt = np.tensordot(np.array(self.Bs), np.array(self.tops),
                              axes=(0, 0))
            self._bt = [[None] * D for _ in range(D)]
            for i in range(D):
                for j in range(i, D):
                    sp = bttb_spectrum(bt[i, j], self.sizes)
                    self._bt[i][j] = self._bt[j][i] = sp
        elif rep == 'slfm':
            # grid_kernel.py:84-112
            self._astar = np.vstack(coreg_vecs).T                  # [D, sum R]
            ranks = [len(a) for a in coreg_vecs]
            self._slfm_specs = [sp for sp, r in zip(self.specs, ranks)
                                for _ in range(r)]
            diags = np.column_stack(coreg_diags)                    # [D, Q]
            diag_tops = diags.dot(np.array(self.tops))
            self._diag_specs = [bttb_spectrum(t, self.sizes)
                                for t in diag_tops]

    def _bttb(self, spec, x):
        return bttb_matvec(spec, self.sizes, x)

    def _grid_sum(self, g, Bs, specs):
        total = 0
        for B, spec in zip(Bs, specs):
            def mm(X, spec=spec):
                return matmat_by_columns(
                    lambda c: self._bttb(spec, c), X, self.m)
            total = total + kron_matvec(B, mm, (self.m, self.m), g)
        return total

    def _grid_bt(self, g):
        # block_matrix.py:32-37
        m, D = self.m, self.D
        out = np.zeros_like(g, dtype=float)
        for i in range(D):
            for j in range(D):
                out[i * m:(i + 1) * m] += self._bttb(
                    self._bt[i][j], g[j * m:(j + 1) * m])
        return out

    def _kron_with_identity(self, A, x):
        # Kronecker(NumpyMatrix(A), Identity(m)).matvec; kronecker.py:39-46
        x = x.reshape(-1, self.m).T
        x = x.reshape(-1, A.shape[1]).T
        x = A.dot(x)
        return x.reshape(-1)

    def _grid_slfm(self, g):
        m = self.m
        if self._coreg_identity:
            coreg = g
        else:
            # Composition([left, toeps, right]) applied right-to-left
            x = self._kron_with_identity(self._astar.T, g)
            y = np.empty(len(self._slfm_specs) * m)
            for r, sp in enumerate(self._slfm_specs):     # block_diag.py:36-40
                y[r * m:(r + 1) * m] = self._bttb(sp, x[r * m:(r + 1) * m])
            coreg = self._kron_with_identity(self._astar, y)
        if self._diag_identity:
            diag = g
        else:
            diag = np.empty(self.D * m)
            for d, sp in enumerate(self._diag_specs):
                diag[d * m:(d + 1) * m] = self._bttb(sp, g[d * m:(d + 1) * m])
        return 0 + coreg + diag

    def grid_matvec(self, g, Bs=None, specs=None):
        if Bs is not None or self.rep == 'sum':
            return self._grid_sum(g, self.Bs if Bs is None else Bs,
                                  self.specs if specs is None else specs)
        if self.rep == 'bt':
            return self._grid_bt(g)
        return self._grid_slfm(g)

    def ski_matvec(self, x, Bs=None, specs=None):
        g = self.WT.dot(x)
        g = self.grid_matvec(g, Bs, specs)
        return self.W.dot(g)

    def matvec(self, x):
        x = np.asarray(x, dtype=float)
        return self.ski_matvec(x) + x * self.noise_rep

    def dense(self):
        return matmat_by_columns(self.matvec, np.identity(self.n), self.n)

    def diagonal(self):
        """diag(K~) without forming K~: K_ii = noise_d + sum_q B_q[d, d] w_i' T_q w_i with w_i the
        (at most 4^d) interpolation weights of point i.  Test oracle of the Jacobi preconditioner
        (not a reference function: the reference only forwards K.preconditioner, iterative.py:47)."""
        W = self.W.tocsr()
        out = np.array(self.noise_rep, dtype=float)
        starts = np.concatenate([[0], np.cumsum(self.lens)])
        sizes = np.array(self.sizes)
        for i in range(self.n):
            d = int(np.searchsorted(starts, i, side='right') - 1)
            cols = W.indices[W.indptr[i]:W.indptr[i + 1]] - d * self.m
            w = W.data[W.indptr[i]:W.indptr[i + 1]]
            idx = np.array(np.unravel_index(cols, sizes)).T             # [nnz, ndim]
            off = np.abs(idx[:, None, :] - idx[None, :, :])
            flat = np.ravel_multi_index(tuple(off[..., a] for a in range(len(sizes))), sizes)
            ww = np.outer(w, w)
            for B, top in zip(self.Bs, self.tops):
                out[i] += B[d, d] * np.sum(ww * top[flat])
        return out


def build_operator(spec, Xs, grids, rep='auto', slfm_identity=(False, False)):
    """gen_grid_kernel (grid_kernel.py:49-74) for one active-dim group; grid
    distances as interpolated_llgp.py:415-431.  ``slfm_identity``: see LmcOperator."""
    W = multi_interpolant_csr(Xs, *grids)
    WT = W.transpose().tocsr()
    mesh = np.stack(np.meshgrid(*grids, indexing='ij'), axis=-1)
    first = mesh.reshape(-1, len(grids))[0]
    dists = np.linalg.norm(mesh - first, axis=-1)
    lens = [len(X) for X in Xs]
    op = LmcOperator(W, WT, spec.coreg_mats(), spec.tops(dists),
                     dists.shape, spec.noise, lens, spec.coreg_vecs,
                     spec.coreg_diags, rep, slfm_identity)
    op.dists = dists
    return op


# ---------------------------------------------------------------------------
# MINRES (scipy/sparse/linalg/_isolve/minres.py, scipy 1.18.1, loop at
# lines 210-364) with M = I, shift = 0, x0 = 0 -- restated, and the
# reference wrapper runlmc/approx/iterative.py:24-62
# ---------------------------------------------------------------------------

def minres(matvec, b, rtol, maxiter, callback=None, trace_at=(), psolve=None, record=None):
    """Paige-Saunders MINRES, arithmetic order as scipy's.

    ``record``: optional list; receives (alfa_k, beta_{k+1}) of every iteration, i.e. the Lanczos
    tridiagonal scipy's recurrence builds (minres.py:247-262), preceded by beta1 as record[0].

    ``psolve``: the preconditioner M (an SPD approximation of the INVERSE, scipy's ``M``:
    ``y = psolve(r)``, minres.py:256 and :313), None = identity.  istop = 9 stands for scipy's
    ValueError on an indefinite preconditioner (r . M r < 0).

    Returns (x, istop, itn, trace) where trace maps iteration -> copy of x for
    the iterations listed in ``trace_at``."""
    b = np.asarray(b, dtype=float)
    n = len(b)
    x = np.zeros(n)
    trace = {}
    r1 = b.copy()
    y = r1 if psolve is None else psolve(r1)
    beta1 = np.inner(r1, y)
    if beta1 < 0:
        return x, 9, 0, trace
    if beta1 == 0:
        return x, 0, 0, trace
    beta1 = math.sqrt(beta1)
    if record is not None:
        record.append(beta1)
    oldb = 0.0
    beta = beta1
    dbar = 0.0
    epsln = 0.0
    phibar = beta1
    rhs1 = beta1
    rhs2 = 0.0
    tnorm2 = 0.0
    gmax = 0.0
    gmin = np.finfo(np.float64).max
    cs = -1.0
    sn = 0.0
    w = np.zeros(n)
    w2 = np.zeros(n)
    r2 = r1
    istop = 0
    itn = 0
    while itn < maxiter:
        itn += 1
        s = 1.0 / beta
        v = s * y
        y = matvec(v)
        if itn >= 2:
            y = y - (beta / oldb) * r1
        alfa = np.inner(v, y)
        y = y - (alfa / beta) * r2
        r1 = r2
        r2 = y
        if psolve is not None:
            y = psolve(r2)
        oldb = beta
        beta = np.inner(r2, y)
        if beta < 0:
            return x, 9, itn, trace
        beta = math.sqrt(beta)
        if record is not None:
            record.append((alfa, beta))
        tnorm2 += alfa ** 2 + oldb ** 2 + beta ** 2
        if itn == 1 and beta / beta1 <= 10 * EPS:
            istop = -1
        oldeps = epsln
        delta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        epsln = sn * beta
        dbar = -cs * beta
        root = np.linalg.norm([gbar, dbar])
        gamma = np.linalg.norm([gbar, beta])
        gamma = max(gamma, EPS)
        cs = gbar / gamma
        sn = beta / gamma
        phi = cs * phibar
        phibar = sn * phibar
        denom = 1.0 / gamma
        w1 = w2
        w2 = w
        w = (v - oldeps * w1 - delta * w2) * denom
        x = x + phi * w
        gmax = max(gmax, gamma)
        gmin = min(gmin, gamma)
        z = rhs1 / gamma
        rhs1 = rhs2 - delta * z
        rhs2 = -epsln * z
        Anorm = math.sqrt(tnorm2)
        ynorm = np.linalg.norm(x)
        epsx = Anorm * ynorm * EPS
        rnorm = phibar
        test1 = np.inf if (ynorm == 0 or Anorm == 0) else rnorm / (Anorm * ynorm)
        test2 = np.inf if Anorm == 0 else root / Anorm
        Acond = gmax / gmin
        if istop == 0:
            if 1 + test2 <= 1:
                istop = 2
            if 1 + test1 <= 1:
                istop = 1
            if itn >= maxiter:
                istop = 6
            if Acond >= 0.1 / EPS:
                istop = 4
            if epsx >= beta1:
                istop = 3
            if test2 <= rtol:
                istop = 2
            if test1 <= rtol:
                istop = 1
        if itn in trace_at:
            trace[itn] = x.copy()
        if callback is not None:
            callback(x)
        if istop != 0:
            break
    return x, istop, itn, trace


def lanczos_quadrature_logdet(beta1, coeffs):
    """z^T log(A) z ~= beta1^2 e1^T log(T_k) e1 for the Lanczos tridiagonal T_k started at z / ||z||
    (stochastic Lanczos quadrature, Ubaru-Chen-Saad 2017, the "Lanczos" item of the reference's roadmap,
    README.md:88-93).  ``coeffs``: [(alfa_j, beta_{j+1})] as recorded by ``minres``."""
    import scipy.linalg
    k = len(coeffs)
    if k == 0:
        return 0.0
    d = np.array([c[0] for c in coeffs])
    e = np.array([c[1] for c in coeffs[:-1]])
    theta, vecs = scipy.linalg.eigh_tridiagonal(d, e)
    return float(beta1 ** 2 * np.sum(vecs[0] ** 2 * np.log(theta)))


def stochastic_logdet(matvec, probes, rtol, maxiter):
    """log det A ~= mean_i z_i^T log(A) z_i over Rademacher probes, each term from the Lanczos process
    of the MINRES solve A x = z_i.  Returns (estimate, per-probe terms)."""
    terms = []
    for z in probes:
        rec = []
        minres(matvec, z, rtol, maxiter, record=rec)
        terms.append(lanczos_quadrature_logdet(rec[0], rec[1:]))
    return float(np.mean(terms)), np.array(terms)


def cg(matvec, b, rtol, maxiter, callback=None):
    """scipy.sparse.linalg.cg (scipy 1.18.1 _isolve/iterative.py, the loop after
    ``make_system``) with M = I, x0 = 0, atol = 0 -- the solver behind
    Iterative.solve(..., minres=False) (iterative.py:44-51).
    Returns (x, info, iterations)."""
    b = np.asarray(b, dtype=float)
    bnrm2 = np.linalg.norm(b)
    atol = max(0.0, float(rtol) * float(bnrm2))
    if bnrm2 == 0:
        return b.copy(), 0, 0
    x = np.zeros_like(b)
    r = b.copy()
    rho_prev, p = None, None
    for iteration in range(maxiter):
        if np.linalg.norm(r) < atol:
            return x, 0, iteration
        z = r                       # identity preconditioner
        rho_cur = np.dot(r, z)
        if iteration > 0:
            beta = rho_cur / rho_prev
            p *= beta
            p += z
        else:
            p = z.copy()
        q = matvec(p)
        alpha = rho_cur / np.dot(p, q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho_cur
        if callback is not None:
            callback(x)
    return x, maxiter, maxiter


class _Early(Exception):
    def __init__(self, x):
        super().__init__('')
        self.x = x


def iterative_solve(matvec, y, tol=1e-4, use_scipy=False, check_every=100, use_minres=True, psolve=None):
    """Iterative.solve with verbose=True (minres=True unless ``use_minres`` is false: then scipy's cg
    restated above runs behind the same wrapper); iterative.py:24-62.

    rtol = min(1e-10, tol), maxiter = n, and on every 100th callback the true
    residual ||y - Kx||_2 is formed; < tol terminates (iterative.py:36-42).
    Returns (x, callbacks, final abs residual).  ``use_scipy`` swaps the
    restated loop for scipy's own minres (same result; used to pin the loop).
    ``psolve``: K.preconditioner, forwarded as scipy's M (iterative.py:47-50; MINRES only)."""
    y = np.asarray(y, dtype=float)
    n = len(y)
    ctr = 0

    def cb(x):
        nonlocal ctr
        ctr += 1
        if ctr % check_every == 0:
            if np.linalg.norm(y - matvec(x)) < tol:
                raise _Early(x)

    rtol = min(1e-10, tol)
    try:
        if use_scipy:
            op = scipy.sparse.linalg.LinearOperator(
                (n, n), matvec=matvec, dtype=np.float64)
            method = scipy.sparse.linalg.minres if use_minres else scipy.sparse.linalg.cg
            M = None if psolve is None else scipy.sparse.linalg.LinearOperator(
                (n, n), matvec=psolve, dtype=np.float64)
            x, _ = method(op, y, rtol=rtol, maxiter=n, M=M, callback=cb)
        elif use_minres:
            x, _, _, _ = minres(matvec, y, rtol, n, callback=cb, psolve=psolve)
        else:
            x, _, _ = cg(matvec, y, rtol, n, callback=cb)
    except _Early as e:
        x = e.x
    err = np.linalg.norm(y - matvec(x))
    return x, ctr, err


# ---------------------------------------------------------------------------
# Stochastic gradient (runlmc/lmc/stochastic_deriv.py, likelihood.py,
# derivative.py)
# ---------------------------------------------------------------------------

def solve_all(op, y, rs, tol=1e-4, pool=None):
    """StochasticDerivService.generate's solves; stochastic_deriv.py:33-52.
    ``rs`` is supplied by the caller (host-side probes) instead of being drawn
    from the global numpy RNG (stochastic_deriv.py:35)."""
    rhs = [y] + [r for r in rs]
    if pool is None:
        sols = [iterative_solve(op.matvec, b, tol) for b in rhs]
    else:
        sols = pool.starmap(_solve_task, [(op, b, tol) for b in rhs])
    alpha = sols[0][0]
    inv_rs = [s[0] for s in sols[1:]]
    iters = [s[1] for s in sols]
    errs = [s[2] for s in sols]
    return alpha, inv_rs, iters, errs


def _solve_task(op, b, tol):
    return iterative_solve(op.matvec, b, tol)


def derivative(dK_matvec, alpha, rs, inv_rs):
    """0.5 (alpha' dK alpha - mean_i inv_r_i' dK r_i);
    derivative.py:5-6, stochastic_deriv.py:69-78."""
    quad = alpha.dot(dK_matvec(alpha))
    trace = 0
    for r, rinv in zip(rs, inv_rs):
        trace += rinv.dot(dK_matvec(r))
    trace = trace / len(rs)
    return 0.5 * (quad - trace)


def gradients(op, spec, alpha, rs, inv_rs):
    """All four gradient families of ApproxLMCLikelihood;
    likelihood.py:48-96 (loops), 112-128 (dK operators).

    Returns dict with 'coreg_vec' (list of [R_q, D]), 'coreg_diag' (list of
    [D]), 'kernel' (list of lists) and 'noise' ([D])."""
    D, Q = spec.D, spec.Q
    sizes = op.sizes
    out = {'coreg_vec': [], 'coreg_diag': [], 'kernel': [], 'noise': None}

    def ski_with(B, spec_q):
        return lambda x: op.ski_matvec(x, [B], [spec_q])

    for q, a in enumerate(spec.coreg_vecs):
        g = np.zeros(a.shape)
        for i, ai in enumerate(a):
            for j in range(D):
                dA = np.zeros((D, D))
                dA[j] += ai
                dA.T[j] += ai
                g[i, j] = derivative(ski_with(dA, op.specs[q]),
                                     alpha, rs, inv_rs)
        out['coreg_vec'].append(g)
    for q in range(Q):
        g = np.zeros(D)
        for i in range(D):
            E = np.zeros((D, D))
            E[i, i] = 1
            g[i] = derivative(ski_with(E, op.specs[q]), alpha, rs, inv_rs)
        out['coreg_diag'].append(g)
    Bs = spec.coreg_mats()
    for q, dtops in enumerate(spec.top_grads(op.dists)):
        gq = []
        for dt in dtops:
            sp = bttb_spectrum(dt.ravel(), sizes)
            gq.append(derivative(ski_with(Bs[q], sp), alpha, rs, inv_rs))
        out['kernel'].append(gq)
    g = np.zeros(D)
    for i in range(D):
        e = np.zeros(D)
        e[i] = 1
        rep = np.repeat(e, op.lens)
        g[i] = derivative(lambda x, rep=rep: x * rep, alpha, rs, inv_rs)
    out['noise'] = g
    return out


def flatten_gradients(g):
    parts = [np.ravel(a) for a in g['coreg_vec']]
    parts += [np.ravel(a) for a in g['coreg_diag']]
    parts += [np.asarray(k, dtype=float) for k in g['kernel']]
    parts.append(np.ravel(g['noise']))
    return np.concatenate(parts)


# ---------------------------------------------------------------------------
# Prediction (runlmc/models/interpolated_llgp.py:293-397) -- the model class
# itself needs paramz, so its arithmetic is restated here on the oracle
# operator; one active-dimension group (the fused path's case).
# ---------------------------------------------------------------------------

def exact_cross_kernel(spec, Xs, Zs):
    """ExactLMCLikelihood.kernel_from_indices (lmc/likelihood.py:183-203):
    dense K[i, j] = sum_q B_q[out(i), out(j)] k_q(|x_i - z_j|), no noise."""
    import scipy.spatial.distance as dist
    rlens, clens = [len(X) for X in Xs], [len(Z) for Z in Zs]
    X = np.vstack([np.asarray(x, dtype=float).reshape(len(x), -1) for x in Xs])
    Z = np.vstack([np.asarray(z, dtype=float).reshape(len(z), -1) for z in Zs])
    r = dist.cdist(X, Z)
    rb = np.concatenate([[0], np.cumsum(rlens)])
    cb = np.concatenate([[0], np.cumsum(clens)])
    K = np.zeros(r.shape)
    for B, kind, kp in zip(spec.coreg_mats(), spec.kinds, spec.kparams):
        Kq = kern_eval(kind, kp, r)
        for i in range(len(rlens)):
            for j in range(len(clens)):
                Kq[rb[i]:rb[i + 1], cb[j]:cb[j + 1]] *= B[i, j]
        K += Kq
    return K


def native_variance(spec):
    """_native_variance (interpolated_llgp.py:304-316): prior variance of one
    point of each output, sum_q (|a_q[:, d]|^2 + kappa_q[d]) k_q(0) + noise_d."""
    coregs = np.column_stack([np.square(np.atleast_2d(a)).sum(axis=0)
                              for a in spec.coreg_vecs])
    coregs = coregs + np.column_stack(spec.coreg_diags)
    k0 = np.array([kern_eval(kind, kp, np.zeros(1))[0]
                   for kind, kp in zip(spec.kinds, spec.kparams)])
    return coregs.dot(k0).reshape(-1) + np.asarray(spec.noise)


def predict_mean(op, alpha, Xs_test, grids):
    """W* (K_UU W^T alpha); _grid_alpha + _raw_predict (interpolated_llgp.py:293-300, 334-338)."""
    grid_alpha = op.grid_matvec(op.WT.dot(alpha))
    Wt = multi_interpolant_csr(Xs_test, *grids)
    return Wt.dot(grid_alpha)


def precomputed_nu(op, tol=1e-4):
    """_precomputed_nu / _var_solve (interpolated_llgp.py:350-388): for every grid
    index i, nu_i = e_i' K_UX K^-1 K_XU e_i with K_XU = W K_UU."""
    Dm = op.W.shape[1]
    nu = np.zeros(Dm)
    for i in range(Dm):
        e = np.zeros(Dm)
        e[i] = 1
        x = op.W.dot(op.grid_matvec(e))
        x, _, _ = iterative_solve(op.matvec, x, tol)
        nu[i] = op.grid_matvec(op.WT.dot(x))[i]
    return nu


def predict_var_precompute(op, Xs_test, grids, nu):
    """_var_predict_precompute (interpolated_llgp.py:384-388): W* nu."""
    return multi_interpolant_csr(Xs_test, *grids).dot(nu)


def predict_var_on_the_fly(op, spec, Xs_train, Xs_test, tol=1e-4):
    """_var_predict_on_the_fly (interpolated_llgp.py:390-397): one solve per test
    point with the exact cross-covariance row as right-hand side."""
    Kx = exact_cross_kernel(spec, Xs_test, Xs_train)
    inv = np.array([iterative_solve(op.matvec, k, tol)[0] for k in Kx]).T
    return np.einsum('ij,ji->i', Kx, inv)


def predict(op, spec, alpha, Xs_train, Xs_test, grids, mode='on-the-fly', tol=1e-4, nu=None):
    """_raw_predict (interpolated_llgp.py:324-348): (mean, var), var clipped at 0."""
    lens = [len(X) for X in Xs_test]
    mean = predict_mean(op, alpha, Xs_test, grids)
    nat = np.repeat(native_variance(spec), lens)
    if mode == 'precompute':
        expl = predict_var_precompute(op, Xs_test, grids, precomputed_nu(op, tol) if nu is None else nu)
    else:
        expl = predict_var_on_the_fly(op, spec, Xs_train, Xs_test, tol)
    var = nat - expl
    var[var < 0] = 0
    return mean, var
