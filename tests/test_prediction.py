"""Prediction paths (reference models/interpolated_llgp.py:293-397):
the oracle restatement against golden vectors produced by the reference's own building blocks
(CPU), and the device Predictor against the same vectors (GPU)."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import lmc_oracle as orc
from runlmc_b200 import synthetic

PREDICT_PROBLEMS = {
    'predict_A': lambda: synthetic.make_problem('A', seed=1234, cells_per_lengthscale=4, grid=[40]),
    'predict_2d': lambda: synthetic.make_problem('e_small', seed=99, cells_per_lengthscale=3,
                                                 lens=[60, 50, 55], grid=[8, 7], N=5),
}


def _xtest(g, D):
    return [g['Xtest_%d' % d] for d in range(D)]


def _oracle(prob):
    spec = orc.KernelSpec(['rbf'] * prob.Q, [[gm] for gm in prob.gammas], prob.coreg_vecs,
                          prob.coreg_diags, prob.noise)
    return spec, orc.build_operator(spec, prob.Xs, prob.grids, rep='sum')


@pytest.mark.parametrize('name', sorted(PREDICT_PROBLEMS))
def test_oracle_prediction_golden(name):
    g = load_golden(name)
    prob = PREDICT_PROBLEMS[name]()
    spec, op = _oracle(prob)
    Xt = _xtest(g, prob.D)
    np.testing.assert_allclose(orc.exact_cross_kernel(spec, Xt, prob.Xs), g['K_test_X'], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(np.repeat(orc.native_variance(spec), g['lens']), g['native'], rtol=1e-14)
    alpha, _, _ = orc.iterative_solve(op.matvec, prob.y, 1e-4)
    assert rel_err(alpha, g['alpha']) < 1e-6      # converged MINRES iterates agree to ~1e-7 (Lanczos)
    assert rel_err(orc.predict_mean(op, g['alpha'], Xt, prob.grids), g['mean']) < 1e-12
    mean, var = orc.predict(op, spec, g['alpha'], prob.Xs, Xt, prob.grids, mode='on-the-fly')
    assert rel_err(var, g['var_fly']) < 1e-6
    # precompute: the D*m solves are slow in pure Python -- check the first grid entries only,
    # and the W* nu product with the stored nu
    assert rel_err(orc.predict_var_precompute(op, Xt, prob.grids, g['nu']),
                   g['native'] - g['var_pre']) < 1e-10 or np.any(g['var_pre'] == 0)
    e = np.zeros(op.W.shape[1])
    e[3] = 1
    x, _, _ = orc.iterative_solve(op.matvec, op.W.dot(op.grid_matvec(e)), 1e-4)
    assert abs(op.grid_matvec(op.WT.dot(x))[3] - g['nu'][3]) < 1e-6 * max(1.0, abs(g['nu'][3]))


def _predictor(prob, alpha):
    from runlmc_b200.fused import FusedLMC
    from runlmc_b200.kern import RBF
    from runlmc_b200.lmc.functional_kernel import FunctionalKernel
    from runlmc_b200.lmc.prediction import Predictor
    fk = FunctionalKernel(D=prob.D, lmc_kernels=[RBF(gm) for gm in prob.gammas], lmc_ranks=[1] * prob.Q)
    fk.noise = prob.noise
    fk.coreg_vecs = prob.coreg_vecs
    fk.coreg_diags = prob.coreg_diags
    fk.set_input_dim(prob.ndim)
    op = FusedLMC(prob.Xs, prob.grids)
    op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    return Predictor(op, fk, prob.Xs, prob.grids, alpha, tol=1e-4, block=32)


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(PREDICT_PROBLEMS))
def test_device_prediction_golden(name):
    from runlmc_b200.lmc.prediction import kernel_from_indices
    g = load_golden(name)
    prob = PREDICT_PROBLEMS[name]()
    pr = _predictor(prob, g['alpha'])
    Xt = _xtest(g, prob.D)
    np.testing.assert_allclose(kernel_from_indices(Xt, prob.Xs, pr.fk), g['K_test_X'], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(np.repeat(pr.native_variance(), g['lens']), g['native'], rtol=1e-13)
    assert rel_err(pr.mean(Xt), g['mean']) < 1e-10
    # the solves stop at the solver tolerance (absolute residual 1e-4), so variances agree to that
    scale = float(np.max(g['native']))
    means, vars_fly = pr.predict(Xt, mode='on-the-fly')
    assert rel_err(np.hstack(means), g['mean']) < 1e-10
    assert np.max(np.abs(np.hstack(vars_fly) - g['var_fly'])) < 1e-3 * scale
    assert np.max(np.abs(pr.nu() - g['nu'])) < 1e-3 * max(1.0, float(np.max(np.abs(g['nu']))))
    _, vars_pre = pr.predict(Xt, mode='precompute')
    assert np.max(np.abs(np.hstack(vars_pre) - g['var_pre'])) < 1e-3 * scale
    assert [len(v) for v in vars_pre] == list(g['lens'])
    with pytest.raises(ValueError):
        pr.predict(Xt, mode='exact')
