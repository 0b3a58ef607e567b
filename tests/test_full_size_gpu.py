"""Properties of the CUDA path at BASELINE.json's full sizes (config D: n = 500k, 1-D grid 8192;
config E: n = 1M, 2-D grid 256 x 256), where the oracle would take minutes per product: linearity,
symmetry, adjointness of the two interpolation stages, agreement of wide blocks with single columns
and of the sorted-order entry point with the caller-order one, and the solver's reported residual
against a residual formed from a separate product."""
import numpy as np
import pytest

from conftest import rel_err
from runlmc_b200 import synthetic

pytestmark = pytest.mark.gpu

CPL = {'D': 2, 'E': 1.5}     # bench.py's lengthscales for these workloads


@pytest.fixture(scope='module', params=['D', 'E'])
def full(request):
    import torch
    from runlmc_b200 import kern
    from runlmc_b200.fused import FusedLMC
    name = request.param
    prob = synthetic.make_problem(name, seed=1234, cells_per_lengthscale=CPL[name], N=2)
    op = FusedLMC(prob.Xs, prob.grids)
    op.set_kernels([kern.RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs,
                   prob.coreg_diags)
    gen = torch.Generator(device='cuda').manual_seed(11)
    P = 129 if name == 'E' else 65
    V = torch.randn(P, prob.n, dtype=torch.float64, device='cuda', generator=gen)
    return name, prob, op, V


def _rel(a, b):
    import torch
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


def test_full_size_linearity_symmetry_definiteness(full):
    import torch
    name, prob, op, V = full
    KV = op.mvm_device(V)
    a, b = 0.7, -1.3
    lin = op.mvm_device((a * V[0] + b * V[1]).reshape(1, -1).contiguous())[0]
    assert _rel(lin, a * KV[0] + b * KV[1]) < 1e-12
    s1, s2 = float(torch.dot(V[2], KV[3])), float(torch.dot(V[3], KV[2]))
    assert abs(s1 - s2) <= 1e-10 * max(abs(s1), abs(float(torch.dot(V[2], KV[2]))))
    assert float(torch.dot(V[4], KV[4])) > 0


def test_full_size_wide_block_equals_single_columns(full):
    import torch
    name, prob, op, V = full
    KV = op.mvm_device(V)
    for c in (0, 1, V.shape[0] // 2, V.shape[0] - 1):      # first pair, a middle column, the odd last column
        one = op.mvm_device(V[c].reshape(1, -1).contiguous())[0]
        assert _rel(KV[c], one) < 1e-13                    # a column sees its pair partner only at rounding level
    again = op.mvm_device(V)
    assert torch.equal(KV, again)                          # deterministic


def test_full_size_sorted_entry_point(full):
    import torch
    name, prob, op, V = full
    perm = torch.as_tensor(op.perm().astype(np.int64), device='cuda')
    KV = op.mvm_device(V)
    KVs = op.mvm_sorted_device(V[:, perm].contiguous())
    assert torch.equal(KVs, KV[:, perm])                   # same kernels, same summation order


def test_full_size_interpolation_stages_are_adjoint(full):
    import torch
    name, prob, op, V = full
    v = V[:3].contiguous()
    gen = torch.Generator(device='cuda').manual_seed(5)
    g = torch.randn(3, prob.D * op.m, dtype=torch.float64, device='cuda', generator=gen)
    WTv = op.to_grid_device(v)
    Wg = op.from_grid_device(g)
    for i in range(3):
        lhs, rhs = float(torch.dot(Wg[i], v[i])), float(torch.dot(g[i], WTv[i]))
        assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), 1.0)
    # partition of unity: W 1 = 1 at every point strictly inside the grid (weights sum to one)
    ones = torch.ones(1, prob.D * op.m, dtype=torch.float64, device='cuda')
    assert float((op.from_grid_device(ones) - 1).abs().max()) < 1e-12
    # grid operator is symmetric too
    Kg = op.grid_mvm_device(g)
    s1, s2 = float(torch.dot(g[0], Kg[1])), float(torch.dot(g[1], Kg[0]))
    assert abs(s1 - s2) <= 1e-10 * max(abs(s1), abs(float(torch.dot(g[0], Kg[0]))))


def test_full_size_solver_residuals(full):
    """40 MINRES / CG iterations on a few columns: the residual each solver reports is the true
    residual of the iterate it returns, and MINRES's decreases monotonically with the budget."""
    import torch
    name, prob, op, V = full
    B = V[:5].contiguous()
    prev = None
    for k in (10, 40):
        X, iters, resid, _ = op.minres_device(B, tol=1e-4, maxiter=k)
        assert (iters == k).all()
        true = torch.linalg.norm(B - op.mvm_device(X), dim=1).cpu().numpy()
        assert np.allclose(resid, true, rtol=1e-9)
        if prev is not None:
            assert (true <= prev * (1 + 1e-12)).all()
        prev = true
    Bh = B.cpu().numpy()
    Xc, itc, rc, info = op.cg(Bh, tol=1e-4, maxiter=25)
    assert (itc == 25).all() and (info == 25).all()
    truec = np.linalg.norm(Bh - op.mvm(Xc), axis=1)
    assert np.allclose(rc, truec, rtol=1e-9)
