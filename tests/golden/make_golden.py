"""Generate golden vectors by running the REAL reference (vlad17/runlmc).

Run in the build container only (the reference checkout is not present on the
GPU box):

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 OMP_NUM_THREADS=1 \
        python tests/golden/make_golden.py

It imports the unmodified reference modules, feeds them seeded inputs and
stores inputs and outputs in ``tests/golden/*.npz``.  The only accommodations
(SURVEY.md section 8c):
  * scipy >= 1.14 renamed minres' ``tol`` to ``rtol``; approx/iterative.py:50
    still passes ``tol=`` -> a keyword shim is installed here.
  * lmc/functional_kernel.py needs paramz (absent); a duck-typed stand-in with
    the surface gen_grid_kernel / ApproxLMCLikelihood consume is used
    (functional_kernel.py:225-300).
Nothing from the reference is copied into this repository; only numbers.
"""
import os
import sys

import numpy as np
import scipy.sparse.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
sys.path.insert(0, '/root/reference')

_orig_minres = sla.minres
sla.minres = lambda A, b, tol=1e-5, **kw: _orig_minres(A, b, rtol=tol, **kw)
_orig_cg = sla.cg
sla.cg = lambda A, b, tol=1e-5, **kw: _orig_cg(A, b, rtol=tol, **kw)   # same rename for the minres=False path

from runlmc.linalg.toeplitz import Toeplitz  # noqa: E402
from runlmc.linalg.bttb import BTTB  # noqa: E402
from runlmc.linalg.kronecker import Kronecker  # noqa: E402
from runlmc.linalg.numpy_matrix import NumpyMatrix  # noqa: E402
from runlmc.linalg.sum_matrix import SumMatrix  # noqa: E402
from runlmc.linalg.diag import Diag  # noqa: E402
from runlmc.approx import interpolation as ref_interp  # noqa: E402
from runlmc.approx.iterative import Iterative  # noqa: E402
from runlmc.lmc.grid_kernel import gen_grid_kernel, GridKernel  # noqa: E402
from runlmc.lmc.likelihood import ApproxLMCLikelihood  # noqa: E402
from runlmc.lmc.stochastic_deriv import StochasticDerivService  # noqa: E402
from runlmc.util.inline_pool import InlinePool  # noqa: E402

from runlmc_b200 import synthetic  # noqa: E402


class DuckKernel:
    """Stand-in for FunctionalKernel; all kernels LMC-type (rank R_q +
    diagonal), one active-dimension group."""

    def __init__(self, prob):
        self.p = prob
        self.D, self.Q = prob.D, prob.Q
        self.noise = prob.noise
        self.coreg_vecs = prob.coreg_vecs
        self.coreg_diags = prob.coreg_diags
        self.ad = tuple(range(prob.ndim))
        self.active_dims = {self.ad: list(range(prob.Q))}
        self.num_lmc = {self.ad: prob.Q}
        self.num_slfm = {self.ad: 0}
        self.num_indep = {self.ad: 0}

    def coreg_mats(self, active_dim=None):
        return self.p.coreg_mats()

    def total_rank(self, active_dim):
        return sum(len(a) for a in self.coreg_vecs)

    def eval_kernels(self, dists):
        return [synthetic.rbf_top(dists[self.ad], g) for g in self.p.gammas]

    def eval_kernels_fixed_dim(self, dists, active_dim):
        return np.array([synthetic.rbf_top(dists, g) for g in self.p.gammas])

    def eval_kernel_gradients(self, dists):
        return [[synthetic.rbf_top_grad(dists[self.ad], g)]
                for g in self.p.gammas]

    def get_active_dims(self, q):
        return self.ad

    def filter_non_indep_idxs(self, idxs):
        return list(idxs)


def linalg_cases():
    out = {}
    rs = np.random.RandomState(7)
    # Toeplitz tops: the deterministic ones of test_toeplitz.py:22-35 plus
    # seeded exponential-decay tops (testing_utils.py:84-94)
    tops = [
        [1.], [1., 0.], [1., 1.], [0., 0.], [1., -1.],
        [3.5] + [0.999] * 5 + [0.] * 110,
        list((np.arange(10) + 1)[::-1].astype(float)),
        list(np.exp(-rs.rand() * np.arange(10))),
        list(np.exp(-rs.rand() * np.arange(50))),
        list(np.exp(-rs.rand() * np.arange(100))),
        list(np.exp(-0.01 * np.arange(300))),
    ]
    for i, t in enumerate(tops):
        t = np.array(t)
        x = np.arange(len(t)) + 1.0
        out['toep_top_%d' % i] = t
        out['toep_x_%d' % i] = x
        out['toep_y_%d' % i] = Toeplitz(t).matvec(x)
    out['toep_count'] = np.array(len(tops))
    # BTTB shapes of test_bttb.py:17-27, arange and random tops, + larger 2-D
    shapes = [(1,), (3,), (2, 3), (10,), (100,), (2, 3, 4), (5, 7), (16, 16),
              (12, 40)]
    k = 0
    for sh in shapes:
        for top in (np.arange(np.prod(sh)).astype(float),
                    rs.rand(int(np.prod(sh)))):
            x = rs.randn(int(np.prod(sh)))
            out['bttb_shape_%d' % k] = np.array(sh)
            out['bttb_top_%d' % k] = top
            out['bttb_x_%d' % k] = x
            out['bttb_y_%d' % k] = BTTB(top, sh).matvec(x)
            k += 1
    out['bttb_count'] = np.array(k)
    # Kronecker(NumpyMatrix, Toeplitz/BTTB) and the SumMatrix(Kronecker...)+Diag
    # shape of test_sum_matrix.py:57-58
    A = rs.randn(3, 3)
    A = A + A.T
    t = np.exp(-0.3 * np.arange(10))
    x = rs.randn(30)
    out['kron_A'], out['kron_top'], out['kron_x'] = A, t, x
    out['kron_y'] = Kronecker(NumpyMatrix(A), Toeplitz(t)).matvec(x)
    mats, As, ts = [], [], []
    for _ in range(5):
        Aq = rs.randn(2, 2)
        Aq = Aq + Aq.T
        tq = np.exp(-rs.rand() * np.arange(5))
        As.append(Aq)
        ts.append(tq)
        mats.append(Kronecker(NumpyMatrix(Aq), Toeplitz(tq)))
    dg = np.ones(10) * 1e-4
    mats.append(Diag(dg))
    x = rs.randn(10)
    out['sum_As'], out['sum_tops'], out['sum_diag'] = (
        np.array(As), np.array(ts), dg)
    out['sum_x'] = x
    out['sum_y'] = SumMatrix(mats).matvec(x)
    return out


def interp_cases():
    out = {}
    rs = np.random.RandomState(11)
    # test_interpolation.py:84-93
    grid = np.arange(-0.1, 10.1, 0.1)
    sample = np.arange(10) + 0.5
    out['c1_grid'], out['c1_sample'] = grid, sample
    out['c1_dense'] = ref_interp.interp_cubic(grid, sample).toarray()
    # clamped / extrapolating samples (test_interpolation.py:66-82)
    grid = np.arange(10.0)
    sample = np.array([-2.5, -2., -1.2, -0.3, 0., 0.4, 1.7, 7.5, 8.2, 8.999,
                       9., 9.6, 10.3, 11., 11.5])
    out['c2_grid'], out['c2_sample'] = grid, sample
    out['c2_dense'] = ref_interp.interp_cubic(grid, sample).toarray()
    # random interior
    grid = np.linspace(0, 1, 50)
    sample = rs.uniform(0, 1, 200)
    out['c3_grid'], out['c3_sample'] = grid, sample
    out['c3_dense'] = ref_interp.interp_cubic(grid, sample).toarray()
    # bicubic: interior, touching and outside the grid
    gx = np.linspace(0, 1, 12)
    gy = np.linspace(-1, 2, 9)
    s2 = np.column_stack([rs.uniform(-0.3, 1.3, 150),
                          rs.uniform(-1.8, 2.8, 150)])
    s2[:5] = [[0, -1], [1, 2], [0.5, 0.5], [-0.2, 2.5], [1.2, -1.5]]
    out['b1_gx'], out['b1_gy'], out['b1_sample'] = gx, gy, s2
    out['b1_dense'] = ref_interp.interp_bicubic(gx, gy, s2).toarray()
    # multi_interpolant 1-D and 2-D (test_interpolation.py:159-185)
    grid = np.arange(-0.1, 10.1, 0.1)
    sa, sb = np.arange(10) + 0.5, np.sin(np.arange(6)) + 4
    out['m1_grid'], out['m1_sa'], out['m1_sb'] = grid, sa, sb
    out['m1_dense'] = ref_interp.multi_interpolant([sa, sb], grid).toarray()
    gx = np.arange(-0.1, 10.1, 0.5)
    gy = np.arange(-2, 11, 0.5)
    sa = np.column_stack([np.arange(10) + 0.5, np.ones(10)])
    sb = np.column_stack([np.sin(np.arange(6)) + 4, np.arange(6.0)])
    out['m2_gx'], out['m2_gy'], out['m2_sa'], out['m2_sb'] = gx, gy, sa, sb
    out['m2_dense'] = ref_interp.multi_interpolant([sa, sb], gx, gy).toarray()
    # autogrid (interpolation.py:179-215)
    Xs = [np.arange(10.0).reshape(-1, 1), np.arange(4.0, 12).reshape(-1, 1)]
    out['ag_a'] = ref_interp.autogrid(Xs, None, None, None)[0]
    out['ag_b'] = ref_interp.autogrid(Xs, [-1], [13], [10])[0]
    return out


def lmc_case(prob, tol=1e-4, n_vec=3, reps=('sum', 'bt', 'slfm'),
             grads=True, seed=5):
    """Full path through the reference: gen_grid_kernel -> K.matvec,
    Iterative.solve, ApproxLMCLikelihood gradients with recorded probes."""
    out = {}
    fk = DuckKernel(prob)
    ad = fk.ad
    W = ref_interp.multi_interpolant(prob.Xs, *prob.grids)
    WT = W.transpose().tocsr()
    dists = {ad: prob.dists}
    interps = {ad: (W, WT)}
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    rs = np.random.RandomState(seed)
    V = rs.randn(n_vec, prob.n)
    out['V'] = V
    out['KV'] = np.array([K.matvec(v) for v in V])
    # the three equivalent grid representations (grid_kernel.py:22-41)
    for rep in reps:
        gk = GridKernel(fk, prob.dists, W, WT, rep, ad)
        out['KV_' + rep] = np.array([gk.matvec(v) for v in V])
    out['WTV'] = np.array([WT.dot(v) for v in V])
    G = rs.randn(n_vec, W.shape[1])
    out['G'] = G
    out['WG'] = np.array([W.dot(g) for g in G])
    gk = GridKernel(fk, prob.dists, W, WT, 'sum', ad)
    out['KUU_G'] = np.array([gk.grid_K.matvec(g) for g in G])
    # solves
    x, ctr, err = Iterative.solve(K, prob.y, verbose=True, minres=True,
                                  tol=tol)
    out['solve_y_x'], out['solve_y_ctr'], out['solve_y_err'] = (
        x, np.array(ctr), np.array(err))
    if grads:
        np.random.seed(seed)
        svc = StochasticDerivService(None, InlinePool(None), prob.N, tol)
        lik = ApproxLMCLikelihood(fk, K, dists, interps, prob.Ys, svc)
        out['probes'] = np.array(lik.deriv._rs, dtype=np.float64)
        out['inv_probes'] = np.array(lik.deriv._inv_rs)
        out['alpha'] = lik.deriv.alpha
        out['g_coreg_vec'] = np.array(lik.coreg_vec_gradients())
        out['g_coreg_diag'] = np.array(lik.coreg_diags_gradients())
        out['g_kernel'] = np.array(lik.kernel_gradients())
        out['g_noise'] = lik.noise_gradient()
    return out


def predict_case(prob, n_test=7, tol=1e-4, seed=3):
    """The three steps of InterpolatedLLGP._raw_predict (models/interpolated_llgp.py:293-397)
    executed with the reference's own building blocks.  The model class itself cannot be imported
    (paramz), so its glue is restated; every number below comes out of reference code:
    multi_interpolant, GridKernel.grid_K.matvec, Composition/Matrix.wrap, Iterative.solve,
    ExactLMCLikelihood.kernel_from_indices."""
    from runlmc.lmc.likelihood import ExactLMCLikelihood
    from runlmc.linalg.composition import Composition
    from runlmc.linalg.matrix import Matrix
    out = {}
    fk = DuckKernel(prob)
    ad = fk.ad
    W = ref_interp.multi_interpolant(prob.Xs, *prob.grids)
    WT = W.transpose().tocsr()
    dists = {ad: prob.dists}
    K, _ = gen_grid_kernel(fk, dists, {ad: (W, WT)}, prob.lens)
    rs = np.random.RandomState(seed)
    Xs_test = [rs.uniform(0.05, 0.95, size=(n_test + d, prob.ndim)) for d in range(prob.D)]
    lens = [len(X) for X in Xs_test]
    alpha = Iterative.solve(K, prob.y, tol=tol)
    # mean: interpolated_llgp.py:293-300, 334-338
    grid_K = K.Ks[0].grid_K
    grid_alpha = grid_K.matvec(WT.dot(alpha))
    Wstar = ref_interp.multi_interpolant(Xs_test, *prob.grids)
    mean = Wstar.dot(grid_alpha)
    # native variance: interpolated_llgp.py:304-316
    coregs = np.column_stack([np.square(a).sum(axis=0) for a in fk.coreg_vecs])
    coregs = coregs + np.column_stack(fk.coreg_diags)
    kernels = np.array([float(np.ravel(k)[0]) for k in fk.eval_kernels({ad: np.zeros(1)})])
    native = np.repeat(coregs.dot(kernels).reshape(-1) + fk.noise, lens)
    # on the fly: interpolated_llgp.py:390-397
    fk.active_dims_list = None
    K_test_X = ExactLMCLikelihood.kernel_from_indices(Xs_test, prob.Xs, _CdistKernel(fk, prob.ndim))
    inverted = np.array([Iterative.solve(K, k, tol=tol) for k in K_test_X]).T
    var_fly = native - np.diag(K_test_X.dot(inverted))
    var_fly[var_fly < 0] = 0
    # precompute: interpolated_llgp.py:350-388
    K_XU = Composition([Matrix.wrap(W.shape, W.dot), grid_K])
    K_UX = Composition([grid_K, Matrix.wrap(WT.shape, WT.dot)])
    Dm = K_XU.shape[1]
    nu = np.zeros(Dm)
    for i in range(Dm):
        x = np.zeros(Dm)
        x[i] = 1
        x = K_XU.matvec(x)
        x = Iterative.solve(K, x, tol=tol)
        nu[i] = K_UX.matvec(x)[i]
    var_pre = native - Wstar.dot(nu)
    var_pre[var_pre < 0] = 0
    out.update(alpha=alpha, mean=mean, native=native, var_fly=var_fly, var_pre=var_pre, nu=nu,
               K_test_X=K_test_X, lens=np.array(lens))
    for d, X in enumerate(Xs_test):
        out['Xtest_%d' % d] = X
    return out


class _CdistKernel:
    """kernel_from_indices indexes inputs as Xs[:, active_dim] (likelihood.py:176-181); the duck
    kernel's active-dim key is a tuple, which numpy treats as one index per axis -- expose the key as
    a list instead so the column selection works for any input dimension."""

    def __init__(self, fk, ndim):
        self.fk, self.D = fk, fk.D
        self.key = _ListKey(range(ndim))
        self.active_dims = [self.key]

    def eval_kernels(self, dists):
        return self.fk.eval_kernels({self.fk.ad: dists[self.key]})

    def coreg_mats(self):
        return self.fk.coreg_mats()


class _ListKey(list):
    def __hash__(self):
        return hash(tuple(self))


class DuckKindKernel(DuckKernel):
    """Stand-in whose kernels are all SLFM-type (rank-1 coregionalisation, no diagonal) or all
    independent GPs (no coregionalisation vectors, diagonal e_d): the cases in which the reference's
    slfm representation substitutes Identity(m) for the empty part (grid_kernel.py:84-86, 101-103)."""

    def __init__(self, prob, kind):
        super().__init__(prob)
        Q, D = prob.Q, prob.D
        self.kind = kind
        if kind == 'slfm':
            self.coreg_diags = [np.zeros(D) for _ in range(Q)]
            self.num_lmc, self.num_slfm = {self.ad: 0}, {self.ad: Q}
        else:
            self.coreg_vecs = [np.zeros((1, D)) for _ in range(Q)]
            self.coreg_diags = [np.eye(D)[q % D] for q in range(Q)]
            self.num_lmc, self.num_indep = {self.ad: 0}, {self.ad: Q}

    def coreg_mats(self, active_dim=None):
        return [a.T.dot(a) + np.diag(k) for a, k in zip(self.coreg_vecs, self.coreg_diags)]

    def total_rank(self, active_dim):
        return 0 if self.kind == 'indep' else sum(len(a) for a in self.coreg_vecs)

    def filter_non_indep_idxs(self, idxs):
        return [] if self.kind == 'indep' else list(idxs)


class DuckGroupsKernel:
    """LMC kernels on TWO active-dimension groups of 2-D inputs: kernels 0 and 1 see input dimension 0,
    kernel 2 sees dimension 1 (gen_grid_kernel loops over fk.active_dims, grid_kernel.py:49-74)."""

    def __init__(self, prob):
        self.p = prob
        self.D, self.Q = prob.D, 3
        self.noise = prob.noise
        self.coreg_vecs = prob.coreg_vecs
        self.coreg_diags = prob.coreg_diags
        self.active_dims = {(0,): [0, 1], (1,): [2]}
        self.num_lmc = {(0,): 2, (1,): 1}
        self.num_slfm = {(0,): 0, (1,): 0}
        self.num_indep = {(0,): 0, (1,): 0}

    def coreg_mats(self, active_dim=None):
        mats = self.p.coreg_mats()
        return mats if active_dim is None else [mats[q] for q in self.active_dims[active_dim]]

    def total_rank(self, active_dim):
        return sum(len(self.coreg_vecs[q]) for q in self.active_dims[active_dim])

    def eval_kernels_fixed_dim(self, dists, active_dim):
        return np.array([synthetic.rbf_top(dists, self.p.gammas[q]) for q in self.active_dims[active_dim]])

    def filter_non_indep_idxs(self, idxs):
        return list(idxs)


def groups_case(out):
    prob = synthetic.make_problem('e_small', seed=31, cells_per_lengthscale=3, lens=[130, 90, 110],
                                  grid=[40, 24], N=5)
    fk = DuckGroupsKernel(prob)
    dists, interps = {}, {}
    for ad, grid in (((0,), prob.grids[0]), ((1,), prob.grids[1])):
        Xs = [X[:, ad[0]] for X in prob.Xs]
        W = ref_interp.multi_interpolant(Xs, grid)
        interps[ad] = (W, W.transpose().tocsr())
        dists[ad] = grid - grid[0]
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    rs = np.random.RandomState(4)
    V = rs.randn(3, prob.n)
    out['groups_V'] = V
    out['groups_KV'] = np.array([K.matvec(v) for v in V])
    x, ctr, err = Iterative.solve(K, prob.y, verbose=True, minres=True, tol=1e-4)
    out['groups_x'], out['groups_ctr'], out['groups_err'] = x, np.array(ctr), np.array(err)
    print('two active-dimension groups: solve callbacks', ctr, 'residual', err)


def main_extra():
    """(1) gen_grid_kernel for SLFM-only and independent-GP-only kernels; (2) Iterative.solve with a
    K.preconditioner (forwarded to scipy as M, iterative.py:47-50): the Jacobi preconditioner
    Diag(1 / diag(K)) with the diagonal taken from the reference's own dense K."""
    out = {}
    prob = synthetic.make_problem('e_small', seed=99, cells_per_lengthscale=3, lens=[120, 90, 100],
                                  grid=[12, 10], N=5)
    W = ref_interp.multi_interpolant(prob.Xs, *prob.grids)
    WT = W.transpose().tocsr()
    rs = np.random.RandomState(3)
    V = rs.randn(3, prob.n)
    out['V'] = V
    for kind in ('slfm', 'indep'):
        fk = DuckKindKernel(prob, kind)
        K, kerns = gen_grid_kernel(fk, {fk.ad: prob.dists}, {fk.ad: (W, WT)}, prob.lens)
        ktype = {'SumMatrix': 'slfm-or-sum', 'SymmSquareBlockMatrix': 'bt'}[type(kerns[fk.ad].grid_K).__name__]
        parts = [type(k).__name__ for k in kerns[fk.ad].grid_K.Ks]
        out[kind + '_parts'] = np.array(parts)
        out[kind + '_KV'] = np.array([K.matvec(v) for v in V])
        out[kind + '_coreg_vecs'] = np.array(fk.coreg_vecs)
        out[kind + '_coreg_diags'] = np.array(fk.coreg_diags)
        print(kind, 'representation', ktype, parts)
    for name, p in (('2d', prob), ('A', synthetic.make_problem('A', seed=1234, edge=True, cells_per_lengthscale=4))):
        fk = DuckKernel(p)
        Wp = ref_interp.multi_interpolant(p.Xs, *p.grids)
        K, _ = gen_grid_kernel(fk, {fk.ad: p.dists}, {fk.ad: (Wp, Wp.transpose().tocsr())}, p.lens)
        dK = np.diag(K.as_numpy()).copy()
        K.preconditioner = Diag(1.0 / dK)
        x, ctr, err = Iterative.solve(K, p.y, verbose=True, minres=True, tol=1e-4)
        out['pre_%s_diag' % name], out['pre_%s_x' % name] = dK, x
        out['pre_%s_ctr' % name], out['pre_%s_err' % name] = np.array(ctr), np.array(err)
        print('preconditioned', name, 'callbacks', ctr, 'residual', err)
    groups_case(out)
    np.savez_compressed(os.path.join(HERE, 'extra.npz'), **out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'extra':
        return main_extra()
    if len(sys.argv) > 1 and sys.argv[1] == 'predict':
        return main_predict()
    if len(sys.argv) > 1 and sys.argv[1] == 'cg':
        return main_cg()
    np.savez_compressed(os.path.join(HERE, 'linalg.npz'), **linalg_cases())
    np.savez_compressed(os.path.join(HERE, 'interp.npz'), **interp_cases())
    # config A (README scale), 1-D, edge-touching inputs
    pa = synthetic.make_problem('A', seed=1234, edge=True,
                                cells_per_lengthscale=4)
    np.savez_compressed(os.path.join(HERE, 'lmc_A.npz'), **lmc_case(pa))
    # small 2-D problem
    pe = synthetic.make_problem('e_small', seed=99, cells_per_lengthscale=3,
                                lens=[120, 90, 100], grid=[12, 10], N=5)
    np.savez_compressed(os.path.join(HERE, 'lmc_2d.npz'), **lmc_case(pe))
    # medium 1-D problem with the bench recipe's own gammas (ill-conditioned:
    # scipy's test1 stop fires before the residual check)
    pb = synthetic.make_problem('B', seed=1234, lens=[400, 380], grid=[128],
                                N=4)
    np.savez_compressed(os.path.join(HERE, 'lmc_B.npz'),
                        **lmc_case(pb, reps=('sum',)))
    print('golden vectors written to', HERE)


def cg_case(prob, tol=1e-4):
    """Iterative.solve(..., minres=False): scipy's cg behind the same wrapper (iterative.py:44-51)."""
    fk = DuckKernel(prob)
    ad = fk.ad
    W = ref_interp.multi_interpolant(prob.Xs, *prob.grids)
    WT = W.transpose().tocsr()
    K, _ = gen_grid_kernel(fk, {ad: prob.dists}, {ad: (W, WT)}, prob.lens)
    x, ctr, err = Iterative.solve(K, prob.y, verbose=True, minres=False, tol=tol)
    return x, np.array(ctr), np.array(err)


def main_cg():
    probs = {
        'lmc_A': synthetic.make_problem('A', seed=1234, edge=True, cells_per_lengthscale=4),
        'lmc_2d': synthetic.make_problem('e_small', seed=99, cells_per_lengthscale=3, lens=[120, 90, 100],
                                         grid=[12, 10], N=5),
        'lmc_B': synthetic.make_problem('B', seed=1234, lens=[400, 380], grid=[128], N=4),
    }
    out = {}
    for name, prob in probs.items():
        out[name + '_x'], out[name + '_ctr'], out[name + '_err'] = cg_case(prob)
        print(name, 'cg callbacks', int(out[name + '_ctr']), 'residual', float(out[name + '_err']))
    np.savez_compressed(os.path.join(HERE, 'cg.npz'), **out)


def main_predict():
    pa = synthetic.make_problem('A', seed=1234, cells_per_lengthscale=4, grid=[40])
    np.savez_compressed(os.path.join(HERE, 'predict_A.npz'), **predict_case(pa))
    pe = synthetic.make_problem('e_small', seed=99, cells_per_lengthscale=3,
                                lens=[60, 50, 55], grid=[8, 7], N=5)
    np.savez_compressed(os.path.join(HERE, 'predict_2d.npz'), **predict_case(pe, n_test=5))
    print('prediction golden vectors written to', HERE)


if __name__ == '__main__':
    main()
