"""GPU tests of the lmc mirror: gen_grid_kernel / GridKernel / SKI /
Iterative.solve / StochasticDerivService / ApproxLMCLikelihood against the
reference's golden outputs, for the fused operator AND the generic
device-backed operator tree (sum / bt / slfm)."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from runlmc_b200 import synthetic

pytestmark = pytest.mark.gpu


def build(prob, fuse=True):
    from runlmc_b200.approx.interpolation import multi_interpolant
    from runlmc_b200.lmc.functional_kernel import FunctionalKernel
    from runlmc_b200.kern import RBF
    fk = FunctionalKernel(D=prob.D, lmc_kernels=[RBF(g) for g in prob.gammas], lmc_ranks=[1] * prob.Q)
    fk.noise = prob.noise
    fk.coreg_vecs = prob.coreg_vecs
    fk.coreg_diags = prob.coreg_diags
    fk.set_input_dim(prob.ndim)
    ad = tuple(range(prob.ndim))
    W = multi_interpolant(prob.Xs, *prob.grids)
    if not fuse:
        W.lmc_geometry = None
    WT = W.transpose().tocsr()
    return fk, {ad: prob.dists}, {ad: (W, WT)}, ad


def golden_problem(name):
    from test_oracle_golden import GOLDEN_PROBLEMS
    return GOLDEN_PROBLEMS[name](), load_golden(name)


@pytest.mark.parametrize('name', ['lmc_A', 'lmc_2d', 'lmc_B'])
@pytest.mark.parametrize('fuse', [True, False])
def test_gen_grid_kernel_matvec(name, fuse):
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel, GridKernel, FusedSumMatrix
    prob, g = golden_problem(name)
    fk, dists, interps, ad = build(prob, fuse)
    K, kerns = gen_grid_kernel(fk, dists, interps, prob.lens)
    assert isinstance(K, FusedSumMatrix) == fuse
    assert K.shape == (prob.n, prob.n)
    for v, kv in zip(g['V'], g['KV']):
        assert rel_err(K.matvec(v), kv) < 1e-10
    assert rel_err(K.matmat(g['V'].T), g['KV'].T) < 1e-10
    assert rel_err(K.matmat(np.ascontiguousarray(g['V'].T)), g['KV'].T) < 1e-10     # C order: point-major path
    if not fuse:
        for rep in ('sum', 'bt', 'slfm'):
            if 'KV_' + rep in g:
                gk = GridKernel(fk, prob.dists, *interps[ad], rep, ad)
                for v, kv in zip(g['V'], g['KV_' + rep]):
                    assert rel_err(gk.matvec(v), kv) < 1e-10
                for gg, kg in zip(g['G'], g['KUU_G']):
                    assert rel_err(gk.grid_K.matvec(gg), kg) < 1e-10


@pytest.mark.parametrize('fuse', [True, False])
def test_iterative_solve(fuse):
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.approx.iterative import Iterative
    prob, g = golden_problem('lmc_B')
    fk, dists, interps, ad = build(prob, fuse)
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    x, ctr, err = Iterative.solve(K, prob.y, verbose=True, minres=True, tol=1e-4)
    assert abs(ctr - int(g['solve_y_ctr'])) <= 3
    assert rel_err(x, g['solve_y_x']) < 1e-5
    assert err <= max(1e-4, 3 * float(g['solve_y_err']))
    x2 = Iterative.solve(K, prob.y)
    np.testing.assert_array_equal(x, x2)       # deterministic


@pytest.mark.parametrize('name', ['lmc_A', 'lmc_2d', 'lmc_B'])
def test_likelihood_gradients_end_to_end(name):
    """Full pipeline with the reference's own probes (seeded global RNG, as
    stochastic_deriv.py:35): solves + gradients vs the reference's output."""
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.lmc.likelihood import ApproxLMCLikelihood
    from runlmc_b200.lmc.stochastic_deriv import StochasticDerivService
    from runlmc_b200.lmc.metrics import Metrics
    from runlmc_b200.util.inline_pool import InlinePool
    prob, g = golden_problem(name)
    fk, dists, interps, ad = build(prob)
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    metrics = Metrics()
    np.random.seed(5)                      # make_golden.py seeds generate() with 5
    svc = StochasticDerivService(metrics, InlinePool(None), prob.N, 1e-4)
    lik = ApproxLMCLikelihood(fk, K, dists, interps, prob.Ys, svc)
    np.testing.assert_array_equal(np.asarray(lik.deriv._rs, dtype=float), g['probes'])
    assert rel_err(lik.alpha(), g['alpha']) < 1e-5
    assert len(metrics.iterations) == 1 and len(metrics.solv_error) == 1
    # solves agree to solver tolerance => gradients agree to a matching level
    assert rel_err(np.array(lik.coreg_vec_gradients()), g['g_coreg_vec']) < 1e-4
    assert rel_err(np.array(lik.coreg_diags_gradients()), g['g_coreg_diag']) < 1e-4
    assert rel_err(np.array(lik.kernel_gradients()), g['g_kernel']) < 1e-4
    assert rel_err(lik.noise_gradient(), g['g_noise']) < 1e-4
    grads = fk.update_gradient(lik)
    assert set(grads) == {'coreg_vecs', 'coreg_diags', 'kernels', 'noise'}


def test_generic_derivative_matches_fused():
    """StochasticDeriv.derivative(dKdt) with arbitrary device-backed dK operators
    (the reference's per-hyper-parameter loop, likelihood.py:48-96) equals the
    fused Gram-matrix path."""
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.lmc.likelihood import ApproxLMCLikelihood, LMCLikelihood
    from runlmc_b200.lmc.stochastic_deriv import StochasticDerivService, StochasticDeriv
    from runlmc_b200.util.inline_pool import InlinePool
    prob, g = golden_problem('lmc_2d')
    fk, dists, interps, ad = build(prob)
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)

    class FixedService(StochasticDerivService):
        def generate(self, K, y, rs=None):
            return StochasticDeriv(g['alpha'], g['probes'], list(g['inv_probes']), prob.N)

    lik = ApproxLMCLikelihood(fk, K, dists, interps, prob.Ys, FixedService(None, InlinePool(None), prob.N, 1e-4))
    fused = (lik.coreg_vec_gradients(), lik.coreg_diags_gradients(), lik.kernel_gradients(), lik.noise_gradient())
    generic = (LMCLikelihood.coreg_vec_gradients(lik), LMCLikelihood.coreg_diags_gradients(lik),
               LMCLikelihood.kernel_gradients(lik), LMCLikelihood.noise_gradient(lik))
    ref = (g['g_coreg_vec'], g['g_coreg_diag'], g['g_kernel'], g['g_noise'])
    for a, b, r in zip(fused, generic, ref):
        assert rel_err(np.array(a), r) < 1e-8
        assert rel_err(np.array(b), r) < 1e-8


@pytest.mark.parametrize('fuse', [True, False])
def test_iterative_solve_cg(fuse):
    """Iterative.solve(..., minres=False) through the fused operator and through the generic
    operator tree (callback products), against the reference's own cg run."""
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.approx.iterative import Iterative
    prob, _ = golden_problem('lmc_B')
    g = load_golden('cg')
    fk, dists, interps, ad = build(prob, fuse)
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    x, ctr, err = Iterative.solve(K, prob.y, verbose=True, minres=False, tol=1e-4)
    assert abs(ctr - int(g['lmc_B_ctr'])) <= 5       # see test_cg_against_reference_golden
    assert rel_err(x, g['lmc_B_x']) < 1e-5
    assert err <= 1e-4


@pytest.mark.parametrize('name', ['A', '2d'])
@pytest.mark.parametrize('path', ['jacobi', 'generic', 'tree'])
def test_preconditioned_solve_matches_reference(name, path):
    """K.preconditioner is forwarded as scipy's M (iterative.py:47-50).  Golden: the reference's own
    Iterative.solve with K.preconditioner = Diag(1 / diag K).  Paths: the fused operator with the
    library's Jacobi preconditioner (lmc_minres_pre), the fused operator with an arbitrary Matrix
    preconditioner and the plain operator tree (both through lmc_minres_generic_pre)."""
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.approx.iterative import Iterative, JacobiPreconditioner, fused_of
    from runlmc_b200.linalg.diag import Diag
    g = load_golden('extra')
    prob, _ = golden_problem('lmc_' + name)
    fk, dists, interps, ad = build(prob, fuse=path != 'tree')
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    dK = g['pre_%s_diag' % name]
    if path == 'jacobi':
        assert rel_err(fused_of(K).diagonal(), dK) < 1e-13
        K.preconditioner = JacobiPreconditioner(K)
        assert rel_err(K.preconditioner.v, 1.0 / dK) < 1e-13
    else:
        K.preconditioner = Diag(1.0 / dK)
    x, ctr, err = Iterative.solve(K, prob.y, verbose=True, tol=1e-4)
    assert abs(ctr - int(g['pre_%s_ctr' % name])) <= 3
    assert rel_err(x, g['pre_%s_x' % name]) < 1e-5
    assert err <= max(1e-4, 3 * float(g['pre_%s_err' % name]))
    # an indefinite M: scipy raises ValueError (minres.py:259), so does the mirror
    K.preconditioner = Diag(-np.ones(prob.n))
    with pytest.raises(ValueError):
        Iterative.solve(K, prob.y)
    K.preconditioner = Diag(1.0 / dK)
    with pytest.raises(NotImplementedError):          # scipy's cg takes M too; the device cg does not
        Iterative.solve(K, prob.y, minres=False)


def test_preconditioned_iterates_match_oracle():
    """Fixed iteration counts of the preconditioned recurrence against the oracle's restated scipy loop."""
    from oracle import lmc_oracle as orc
    from runlmc_b200.fused import FusedLMC
    prob = synthetic.make_problem('d_small', seed=11, cells_per_lengthscale=6)
    op = FusedLMC(prob.Xs, prob.grids)
    op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    spec = orc.KernelSpec(['rbf'] * prob.Q, [[x] for x in prob.gammas], prob.coreg_vecs, prob.coreg_diags,
                          prob.noise)
    ref = orc.build_operator(spec, prob.Xs, prob.grids, rep='sum')
    dK = ref.diagonal()
    assert rel_err(op.diagonal(), dK) < 1e-13
    RHS = np.vstack([prob.y[None, :], prob.probes[:4]])          # 5 columns: two pairs and an odd one
    for k in (1, 2, 3, 7):
        X, iters, _, istop = op.minres(RHS, tol=1e-4, maxiter=k, check_every=10 ** 6, precond='jacobi')
        for b, x, it in zip(RHS, X, iters):
            xr, _, itn_r, _ = orc.minres(ref.matvec, b, 1e-10, k, psolve=lambda r: r / dK)
            assert it == itn_r
            assert rel_err(x, xr) < 1e-10


def test_operators_keep_their_own_parameters():
    """Two operators built from the same interpolants share one device handle; each must keep the
    hyper-parameters it was built with (the reference returns independent operators)."""
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.approx.iterative import fused_of
    prob, g = golden_problem('lmc_2d')
    fk, dists, interps, ad = build(prob)
    K1, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    v = g['V'][0]
    before = K1.matvec(v)
    assert rel_err(before, g['KV'][0]) < 1e-10
    fk.noise = 3.0 * prob.noise                        # in place, like an optimiser step
    fk.coreg_diags = [2.0 * k for k in prob.coreg_diags]
    for k in fk._kernels:
        k.inv_lengthscale = 1.7 * k.inv_lengthscale
    K2, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    assert fused_of(K1) is fused_of(K2)                # one handle ...
    after2 = K2.matvec(v)
    assert rel_err(after2, before) > 1e-2
    np.testing.assert_array_equal(K1.matvec(v), before)      # ... two parameter states
    np.testing.assert_array_equal(K2.matvec(v), after2)
    # and the generic tree of K1 still holds K1's values too
    assert rel_err(K1.Ks[0].matvec(v) + K1.Ks[1].matvec(v), before) < 1e-10


@pytest.mark.parametrize('kind', ['slfm', 'indep'])
def test_slfm_identity_quirk_is_not_fused(kind):
    """SLFM-only / independent-GP-only kernels with Q > 1: the reference's slfm tree carries Identity(D m)
    for the empty part (grid_kernel.py:84-86, 101-103).  The fused operator cannot represent that, so
    gen_grid_kernel must leave these models to the tree -- which reproduces the reference."""
    from runlmc_b200.approx.interpolation import multi_interpolant
    from runlmc_b200.lmc.functional_kernel import FunctionalKernel
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel, FusedSumMatrix
    from runlmc_b200.kern import RBF
    g = load_golden('extra')
    prob, _ = golden_problem('lmc_2d')
    kerns = [RBF(x) for x in prob.gammas]
    if kind == 'slfm':
        fk = FunctionalKernel(D=prob.D, slfm_kernels=kerns)
    else:
        fk = FunctionalKernel(D=prob.D, indep_gp=kerns, indep_gp_index=[q % prob.D for q in range(prob.Q)])
    fk.noise = prob.noise
    fk.coreg_vecs = list(g[kind + '_coreg_vecs'])
    fk.coreg_diags = list(g[kind + '_coreg_diags'])
    fk.set_input_dim(prob.ndim)
    ad = tuple(range(prob.ndim))
    W = multi_interpolant(prob.Xs, *prob.grids)
    K, _ = gen_grid_kernel(fk, {ad: prob.dists}, {ad: (W, W.transpose().tocsr())}, prob.lens)
    assert not isinstance(K, FusedSumMatrix)
    for v, kv in zip(g['V'], g[kind + '_KV']):
        assert rel_err(K.matvec(v), kv) < 1e-10


@pytest.mark.parametrize('fuse', [True, False])
def test_two_active_dimension_groups(fuse):
    """Kernels on two groups of input dimensions (gen_grid_kernel loops over fk.active_dims,
    grid_kernel.py:49-74): one fused device operator per group, or the plain tree; product and solve against
    the reference's own run."""
    from runlmc_b200.approx.interpolation import multi_interpolant
    from runlmc_b200.approx.iterative import Iterative
    from runlmc_b200.lmc.functional_kernel import FunctionalKernel
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel, FusedGroupsSumMatrix
    from runlmc_b200.kern import RBF
    g = load_golden('extra')
    prob = synthetic.make_problem('e_small', seed=31, cells_per_lengthscale=3, lens=[130, 90, 110],
                                  grid=[40, 24], N=5)
    kerns = [RBF(prob.gammas[0], active_dims=[0]), RBF(prob.gammas[1], active_dims=[0]),
             RBF(prob.gammas[2], active_dims=[1])]
    fk = FunctionalKernel(D=prob.D, lmc_kernels=kerns, lmc_ranks=[1, 1, 1])
    fk.noise = prob.noise
    fk.coreg_vecs = prob.coreg_vecs
    fk.coreg_diags = prob.coreg_diags
    fk.set_input_dim(2)
    assert fk.active_dims == {(0,): [0, 1], (1,): [2]}
    dists, interps = {}, {}
    for ad, grid in (((0,), prob.grids[0]), ((1,), prob.grids[1])):
        W = multi_interpolant([X[:, ad[0]] for X in prob.Xs], grid)
        if not fuse:
            W.lmc_geometry = None
        interps[ad] = (W, W.transpose().tocsr())
        dists[ad] = grid - grid[0]
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    assert isinstance(K, FusedGroupsSumMatrix) == fuse
    for v, kv in zip(g['groups_V'], g['groups_KV']):
        assert rel_err(K.matvec(v), kv) < 1e-10
    x, ctr, err = Iterative.solve(K, prob.y, verbose=True, tol=1e-4)
    assert abs(ctr - int(g['groups_ctr'])) <= 3
    assert rel_err(x, g['groups_x']) < 1e-5
    assert err <= max(1e-4, 3 * float(g['groups_err']))


def test_fused_operator_survives_pickling():
    """The reference ships K to pool workers by pickling it (stochastic_deriv.py:51-52).  The device
    handle does not travel; the unpickled operator walks its (device-backed) tree instead."""
    import pickle
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.approx.iterative import fused_of
    prob, g = golden_problem('lmc_A')
    fk, dists, interps, ad = build(prob)
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    K2 = pickle.loads(pickle.dumps(K))
    assert fused_of(K2) is None
    for v, kv in zip(g['V'], g['KV']):
        assert rel_err(K2.matvec(v), kv) < 1e-10


def test_starmap_of_solves_is_one_block_solve():
    """The batching seam (stochastic_deriv.py:39-52): starmap(Iterative.solve, tasks) over one shared K."""
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.approx.iterative import Iterative
    from runlmc_b200.util.inline_pool import InlinePool
    from runlmc_b200 import _native as nat
    prob, g = golden_problem('lmc_B')
    fk, dists, interps, ad = build(prob)
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    tasks = [(K, prob.y, True, True, 1e-4)] + [(K, r, False, True, 1e-4) for r in prob.probes[:3]]
    l0 = nat.lib.lmc_launch_count()
    one = Iterative.solve(*tasks[0])
    per_solve = nat.lib.lmc_launch_count() - l0
    l0 = nat.lib.lmc_launch_count()
    out = InlinePool(None).starmap(Iterative.solve, tasks)
    batched = nat.lib.lmc_launch_count() - l0
    assert batched < 2 * per_solve                       # not four solves
    x, ctr, err = out[0]
    assert abs(ctr - int(g['solve_y_ctr'])) <= 3 and rel_err(x, g['solve_y_x']) < 1e-5
    assert rel_err(x, one[0]) < 1e-5
    for r, xr in zip(prob.probes[:3], out[1:]):
        assert xr.shape == (prob.n,)
        assert np.linalg.norm(K.matvec(xr) - r) < 1e-2
    # anything else runs task by task
    assert InlinePool(None).starmap(lambda a, b: a + b, [(1, 2), (3, 4)]) == [3, 7]


def test_lanczos_record_and_stochastic_logdet():
    """lmc_minres_lanczos: the tridiagonals the block solver records are the oracle's (restated scipy loop) --
    tightly while Lanczos is still reproducible, and as a quadrature at convergence -- and the log-det estimate
    built from them matches the oracle's estimate on the same probes and the dense log det within its own
    standard error."""
    import torch
    import oracle.lmc_oracle as orc
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.approx.iterative import fused_of
    from runlmc_b200.approx.logdet import stochastic_logdet, quadrature_terms
    from test_oracle_golden import oracle_operator
    prob, _ = golden_problem('lmc_A')
    fk, dists, interps, ad = build(prob)
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    _, ref = oracle_operator(prob)
    n = prob.n
    rng = np.random.default_rng(11)
    probes = rng.integers(0, 2, (24, n)) * 2.0 - 1.0
    fused = fused_of(K)
    R = torch.as_tensor(probes, device='cuda')
    # (a) first steps against the oracle's record
    _, iters, _, _, tri, beta1 = fused.minres_lanczos_device(R[:3].contiguous(), 7, tol=1e-4, maxiter=7)
    for c in range(3):
        rec = []
        orc.minres(ref.matvec, probes[c], 1e-10, 7, record=rec)
        assert abs(beta1[c] - rec[0]) < 1e-12 * rec[0]
        want = np.array(rec[1:])
        assert np.abs(tri[c, :len(want)] - want).max() / np.abs(want).max() < 1e-10
    # (b) a record shorter than the solve keeps the first k steps; a longer one is zero past the last iteration
    _, iters, _, _, tri_s, _ = fused.minres_lanczos_device(R[:3].contiguous(), 5, tol=1e-4, maxiter=7)
    assert np.array_equal(tri_s, tri[:, :5])
    _, iters_f, _, _, tri_f, b1_f = fused.minres_lanczos_device(R[:3].contiguous(), 4000, tol=1e-4)
    for c in range(3):
        assert 0 < iters_f[c] < 4000 and not tri_f[c, iters_f[c]:].any() and tri_f[c, iters_f[c] - 1].all()
    # (c) the estimate
    est, err, terms = stochastic_logdet(K, probes=probes)
    want_est, want_terms = orc.stochastic_logdet(ref.matvec, probes, 1e-10, n)
    assert np.abs(terms - want_terms).max() < 1e-6 * np.abs(want_terms).max()
    sign, logdet = np.linalg.slogdet(ref.dense())
    assert sign > 0 and abs(est - logdet) < 4 * err + 1e-3 * abs(logdet)
    assert len(quadrature_terms(tri_f, b1_f, iters_f)) == 3
    with pytest.raises(ValueError):
        stochastic_logdet(object())
    # (d) as a by-product of the derivative service's own probe solves
    from runlmc_b200.lmc.stochastic_deriv import StochasticDerivService
    deriv = StochasticDerivService(None, None, len(probes), 1e-4, logdet_steps=4000).generate(K, prob.y, rs=probes)
    assert abs(deriv.log_det_K - est) < 1e-9 * abs(est) and abs(deriv.log_det_K_stderr - err) < 1e-6 * err
    plain = StochasticDerivService(None, None, len(probes), 1e-4).generate(K, prob.y, rs=probes)
    assert plain.log_det_K is None and np.array_equal(plain.alpha, deriv.alpha)


def test_device_probes_and_device_resident_derivative():
    """The service keeps right-hand sides and solutions on the device (host copies appear only when read),
    and can draw its Rademacher probes there: reproducible under torch's seed, +-1, and an estimator of the
    same gradient (checked against the exact trace term on a small problem)."""
    import torch
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.lmc.likelihood import ApproxLMCLikelihood
    from runlmc_b200.lmc.stochastic_deriv import StochasticDerivService
    from runlmc_b200.util.inline_pool import InlinePool
    from test_oracle_golden import oracle_operator
    prob, g = golden_problem('lmc_A')
    fk, dists, interps, ad = build(prob)
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    N = 64

    def run(seed):
        torch.manual_seed(seed)
        svc = StochasticDerivService(None, InlinePool(None), N, 1e-6, device_probes=True)
        return ApproxLMCLikelihood(fk, K, dists, interps, prob.Ys, svc)

    lik = run(3)
    d = lik.deriv
    assert d._rs_host is None and d._inv_rs_host is None          # nothing copied to the host yet
    noise = lik.noise_gradient()
    again = run(3).noise_gradient()
    np.testing.assert_array_equal(noise, again)                     # same seed, same probes
    assert not np.array_equal(noise, run(4).noise_gradient())
    rs = d._rs
    assert rs.shape == (N, prob.n) and set(np.unique(rs)) == {-1.0, 1.0}
    assert len(d._inv_rs) == N and rel_err(d.alpha, g['alpha']) < 1e-5
    # exact noise gradient from the dense oracle operator: 0.5 (alpha_d' alpha_d - tr_d(K^-1))
    _, ref = oracle_operator(prob)
    Kinv = np.linalg.inv(ref.dense())
    alpha = Kinv.dot(prob.y)
    off = np.cumsum([0] + list(prob.lens))
    exact = np.array([0.5 * (alpha[a:b].dot(alpha[a:b]) - np.trace(Kinv[a:b, a:b])) for a, b in zip(off, off[1:])])
    # Hutchinson error of the trace term: ~ sqrt(2 / N) ||K^-1_dd||_F per output
    slack = np.array([4 * 0.5 * np.sqrt(2.0 / N) * np.linalg.norm(Kinv[a:b, a:b]) for a, b in zip(off, off[1:])])
    assert np.all(np.abs(noise - exact) < slack + 1e-6 * np.abs(exact))
