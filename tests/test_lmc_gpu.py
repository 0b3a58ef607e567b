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


def test_preconditioner_is_refused_loudly():
    """iterative.py:47 forwards K.preconditioner to scipy; the device solvers are M = I only."""
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.approx.iterative import Iterative
    prob, _ = golden_problem('lmc_A')
    fk, dists, interps, ad = build(prob)
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    K.preconditioner = K
    with pytest.raises(NotImplementedError):
        Iterative.solve(K, prob.y)
