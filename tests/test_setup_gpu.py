"""GPU tests of the operator setup that runs on the device:
the point sort of lmc_op_create_dev against the host builder (bit for bit), and the kernel values /
derivative tops of lmc_op_set_kernels against the numpy formulas of the reference's kern classes
(restated in runlmc_b200.kern and oracle.lmc_oracle)."""
import numpy as np
import pytest

from conftest import rel_err
from runlmc_b200 import synthetic

pytestmark = pytest.mark.gpu

TOP_TOL = 1e-13      # device exp/sin/cos vs libm: a few ulp
MVM_TOL = 1e-10


def _both_builds(prob):
    from runlmc_b200.fused import FusedLMC
    ops = []
    for build in ('host', 'device'):
        op = FusedLMC(prob.Xs, prob.grids, build=build)
        op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
        ops.append(op)
    return ops


BUILD_PROBLEMS = {
    'A_edge': lambda: synthetic.make_problem('A', seed=1234, edge=True, cells_per_lengthscale=4),
    '2d_edge': lambda: synthetic.make_problem('e_small', seed=7, cells_per_lengthscale=3, edge=True),
    '2d_rect': lambda: synthetic.make_problem('e_small', seed=8, cells_per_lengthscale=3,
                                              lens=[500, 0, 433], grid=[40, 9], N=3),
    'ragged': lambda: synthetic.make_problem('d_small', seed=12, cells_per_lengthscale=6,
                                             lens=[1, 350, 0, 77], grid=[100], N=3),
    'C': lambda: synthetic.make_problem('C', seed=13, cells_per_lengthscale=5),
    '2d_mid': lambda: synthetic.make_problem('e_small', seed=21, cells_per_lengthscale=2,
                                             lens=[30000, 25000, 28000], grid=[64, 48], N=3),
}


@pytest.mark.parametrize('name', sorted(BUILD_PROBLEMS))
def test_device_point_build_equals_host_build(name):
    prob = BUILD_PROBLEMS[name]()
    op_h, op_d = _both_builds(prob)
    np.testing.assert_array_equal(op_h.perm(), op_d.perm())
    rng = np.random.default_rng(3)
    V = rng.standard_normal((5, prob.n))
    np.testing.assert_array_equal(op_h.mvm(V), op_d.mvm(V))      # same sort => same summation order
    import torch
    Vd = torch.as_tensor(V, device='cuda')
    np.testing.assert_array_equal(op_h.to_grid_device(Vd).cpu().numpy(), op_d.to_grid_device(Vd).cpu().numpy())


def test_device_point_build_out_of_range_and_duplicates():
    """Points outside the grid (clamped stencils), on grid nodes, and many points in one bin."""
    rng = np.random.default_rng(5)
    grids = [np.linspace(0, 1, 33)]
    Xs = [np.concatenate([rng.uniform(-0.2, 1.2, 300), np.full(200, 0.5), grids[0][::4]]),
          np.concatenate([np.full(64, 1.0), np.full(64, 0.0), rng.uniform(0, 1, 100)])]
    prob = synthetic.make_problem('A', seed=2, cells_per_lengthscale=4, grid=[33], lens=[len(x) for x in Xs])
    prob.Xs = Xs
    op_h, op_d = _both_builds(prob)
    np.testing.assert_array_equal(op_h.perm(), op_d.perm())
    V = rng.standard_normal((3, prob.n))
    np.testing.assert_array_equal(op_h.mvm(V), op_d.mvm(V))


def test_device_point_build_identity_order():
    """Inputs already sorted by grid cell: the permutation is the identity in both builders."""
    grids = [np.linspace(0, 1, 64)]
    Xs = [np.sort(np.random.default_rng(1).uniform(0.02, 0.98, 500))]
    prob = synthetic.make_problem('A', seed=2, cells_per_lengthscale=4, grid=[64], lens=[500], D=1)
    prob.Xs = Xs
    op_h, op_d = _both_builds(prob)
    np.testing.assert_array_equal(op_d.perm(), np.arange(500))
    np.testing.assert_array_equal(op_h.perm(), op_d.perm())
    V = np.random.default_rng(2).standard_normal((2, 500))
    np.testing.assert_array_equal(op_h.mvm(V), op_d.mvm(V))


def test_device_point_build_rejects_non_finite():
    from runlmc_b200.fused import FusedLMC
    prob = BUILD_PROBLEMS['A_edge']()
    Xs = [x.copy() for x in prob.Xs]
    Xs[1][7] = np.nan
    for build in ('host', 'device'):
        with pytest.raises(ValueError):
            FusedLMC(Xs, prob.grids, build=build)
    with pytest.raises(ValueError):
        FusedLMC(prob.Xs, prob.grids, build='nowhere')


def _kernels(kind, Q):
    from runlmc_b200 import kern
    g = np.geomspace(40.0, 400.0, Q)
    if kind == 'rbf':
        return [kern.RBF(x) for x in g]
    if kind == 'matern32':
        return [kern.Matern32(np.sqrt(x)) for x in g]
    if kind == 'periodic':
        return [kern.StdPeriodic(x / 40.0, 0.3 + 0.1 * i) for i, x in enumerate(g)]
    return [kern.RBF(g[0]), kern.Matern32(np.sqrt(g[1])), kern.StdPeriodic(2.0, 0.4)][:Q]


@pytest.mark.parametrize('kind', ['rbf', 'matern32', 'periodic', 'mix'])
@pytest.mark.parametrize('base', ['d_small', 'e_small'])
def test_device_kernel_tops_match_numpy(kind, base):
    from runlmc_b200.fused import FusedLMC
    prob = synthetic.make_problem(base, seed=3, cells_per_lengthscale=4)
    kerns = _kernels(kind, prob.Q)
    op = FusedLMC(prob.Xs, prob.grids)
    np.testing.assert_allclose(op.grid_dists(), prob.dists, rtol=1e-14, atol=1e-15)
    op.set_kernels(kerns, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    want = np.array([k.from_dist(prob.dists).ravel() for k in kerns])
    got = op.kernel_tops()
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= TOP_TOL * np.max(np.abs(want))
    wantd = np.array([t.ravel() for k in kerns for t in k.kernel_gradient(prob.dists)])
    gotd = op.kernel_tops(deriv=True)
    assert gotd.shape == wantd.shape
    for a, b in zip(gotd, wantd):
        assert np.max(np.abs(a - b)) <= TOP_TOL * max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize('kind', ['rbf', 'mix'])
@pytest.mark.parametrize('base', ['d_small', 'e_small'])
def test_set_kernels_equals_set_params(kind, base):
    """Product and gradient Gram matrices with device-evaluated tops vs uploaded tops."""
    from runlmc_b200.fused import FusedLMC
    prob = synthetic.make_problem(base, seed=4, cells_per_lengthscale=4)
    kerns = _kernels(kind, prob.Q)
    tops = [k.from_dist(prob.dists) for k in kerns]
    dtops = [t for k in kerns for t in k.kernel_gradient(prob.dists)]
    op_u = FusedLMC(prob.Xs, prob.grids)
    op_u.set_params(tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    op_k = FusedLMC(prob.Xs, prob.grids)
    op_k.set_kernels(kerns, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    assert op_k.kernels_on_device and not op_u.kernels_on_device
    rng = np.random.default_rng(0)
    V = rng.standard_normal((4, prob.n))
    assert rel_err(op_k.mvm(V), op_u.mvm(V)) < 1e-13
    alpha = rng.standard_normal(prob.n)
    R = prob.probes
    RINV = rng.standard_normal(R.shape)
    gu = op_u.grad_grams(alpha, R, RINV, dtops)
    gk = op_k.grad_grams(alpha, R, RINV, None)
    for a, b in zip(gk, gu):
        assert a.shape == b.shape
        assert rel_err(a, b) < 1e-12
    with pytest.raises(ValueError):
        op_u.grad_grams(alpha, R, RINV, None)          # no descriptors after set_params
    op_k.set_params(tops, prob.coreg_mats(), prob.noise)
    with pytest.raises(ValueError):
        op_k.grad_grams(alpha, R, RINV, None)          # set_params forgets them
    from runlmc_b200 import kern
    with pytest.raises(ValueError):
        op_u.set_kernels([kern.StdPeriodic(1.0, -1.0)] * prob.Q, prob.coreg_mats(), prob.noise)


def test_gen_grid_kernel_uses_device_kernels_only_on_matching_distances():
    """gen_grid_kernel switches to device-evaluated kernels when the caller's grid distances are
    the operator's own; foreign distances (or foreign kernel classes) keep the uploaded tops."""
    from test_lmc_gpu import build, golden_problem
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.approx.iterative import fused_of
    prob, g = golden_problem('lmc_2d')
    fk, dists, interps, ad = build(prob)
    K, _ = gen_grid_kernel(fk, dists, interps, prob.lens)
    assert fused_of(K).kernels_on_device
    assert rel_err(K.matvec(g['V'][0]), g['KV'][0]) < MVM_TOL
    warped = {k: v * 1.01 for k, v in dists.items()}
    K2, _ = gen_grid_kernel(fk, warped, interps, prob.lens)
    assert not fused_of(K2).kernels_on_device
    assert rel_err(K2.matvec(g['V'][0]), g['KV'][0]) > 1e-6     # really the warped kernel


def test_sharded_gradient_with_device_kernels():
    """A whole gradient evaluation with device-evaluated kernels and derivative tops (what bench.py
    times) against the same evaluation with uploaded tops."""
    from test_oracle_golden import GOLDEN_PROBLEMS
    from runlmc_b200 import kern
    from runlmc_b200.fused import FusedLMC
    from runlmc_b200.distributed import sharded_gradient
    prob = GOLDEN_PROBLEMS['lmc_2d']()
    op_u = FusedLMC(prob.Xs, prob.grids, build='host')
    op_u.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    op_k = FusedLMC(prob.Xs, prob.grids)
    op_k.set_kernels([kern.RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs,
                     prob.coreg_diags)
    gu, su = sharded_gradient(op_u, prob.y, prob.probes, prob.top_grads, prob.coreg_vecs, prob.coreg_mats())
    gk, sk = sharded_gradient(op_k, prob.y, prob.probes, None, prob.coreg_vecs, prob.coreg_mats())
    assert rel_err(sk['alpha'], su['alpha']) < 1e-6            # tops differ by ulps; MINRES amplifies
    for a, b in zip(gk, gu):
        assert rel_err(np.array(a, dtype=float), np.array(b, dtype=float)) < 1e-5
