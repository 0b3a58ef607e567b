"""The CUDA path at BASELINE.json's full sizes (config B: n = 10k, 1-D grid 1024; config D: n = 500k,
1-D grid 8192; config E: n = 1M, 2-D grid 256 x 256, D = 10 -- the 512-point fused column kernel the
headline number is measured on).

1. Oracle parity at the stated sizes: the product, its three stages and the first MINRES iterates
   against the numpy/scipy oracle on identical inputs (the oracle builds E's 16 M-nnz interpolation
   matrix in a few seconds and takes ~0.25 s per product), <= 1e-10 relative per column.
2. Size-independent properties on full-width blocks: linearity, symmetry, adjointness of the two
   interpolation stages, agreement of wide blocks with single columns and of the sorted-order entry
   point with the caller-order one, and the solver's reported residual against a residual formed
   from a separate product."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import lmc_oracle as orc
from runlmc_b200 import synthetic

pytestmark = pytest.mark.gpu

CPL = {'B': 8, 'D': 2, 'E': 1.5}     # bench.py's lengthscales for these workloads
WIDTH = {'B': 17, 'D': 65, 'E': 129}  # y + the config's probes
MVM_TOL = 1e-10                       # north_star: MVMs within 1e-10 relative


@pytest.fixture(scope='module', params=['B', 'D', 'E'])
def stated(request):
    """Operator, oracle and a full-width block of the config's own right-hand sides (y + probes)."""
    from runlmc_b200 import kern
    from runlmc_b200.fused import FusedLMC
    name = request.param
    prob = synthetic.make_problem(name, seed=1234, cells_per_lengthscale=CPL[name])
    op = FusedLMC(prob.Xs, prob.grids)
    op.set_kernels([kern.RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs,
                   prob.coreg_diags)
    spec = orc.KernelSpec(['rbf'] * prob.Q, [[g] for g in prob.gammas], prob.coreg_vecs,
                          prob.coreg_diags, prob.noise)
    ref = orc.build_operator(spec, prob.Xs, prob.grids, rep='sum')
    V = np.ascontiguousarray(np.vstack([prob.y[None, :], prob.probes]))
    assert V.shape[0] == WIDTH[name]
    return name, prob, op, ref, V


def test_stated_size_product_matches_oracle(stated):
    """Full-width block through the caller-order and the sorted-order entry points; the first pair,
    a middle column and the odd last column against the oracle."""
    import torch
    name, prob, op, ref, V = stated
    Vd = torch.as_tensor(V, device='cuda')
    KV = op.mvm_device(Vd)
    perm = torch.as_tensor(op.perm().astype(np.int64), device='cuda')
    KVs = op.mvm_sorted_device(Vd[:, perm].contiguous())
    cols = [0, 1, V.shape[0] // 2, V.shape[0] - 1]
    for c in cols:
        want = ref.matvec(V[c])
        assert rel_err(KV[c].cpu().numpy(), want) < MVM_TOL
        assert rel_err(KVs[c].cpu().numpy(), want[op.perm()]) < MVM_TOL
    # the same block point-major ([n, P], the layout bench.py times): row staging in the scatter, row-writing gather
    KVr = op.matmat_device(Vd.t().contiguous())
    assert torch.equal(KVr.t(), KV)                       # bit for bit the column-major entry point
    for c in cols:
        assert rel_err(KVr[:, c].cpu().numpy(), ref.matvec(V[c])) < MVM_TOL
    # the host-buffer entry points (chunked copy/compute pipelines) on the same block
    got = op.mvm(V[[0, 1, V.shape[0] - 1]])
    for g, c in zip(got, (0, 1, V.shape[0] - 1)):
        assert rel_err(g, ref.matvec(V[c])) < MVM_TOL
    got = op.matmat(np.ascontiguousarray(V[[0, 1, V.shape[0] - 1]].T))
    for g, c in zip(got.T, (0, 1, V.shape[0] - 1)):
        assert rel_err(g, ref.matvec(V[c])) < MVM_TOL


def test_stated_size_stages_match_oracle(stated):
    import torch
    name, prob, op, ref, V = stated
    rng = np.random.default_rng(17)
    v = V[[0, 1, V.shape[0] - 1]]
    g = rng.standard_normal((3, prob.D * ref.m))
    vd, gd = torch.as_tensor(v, device='cuda'), torch.as_tensor(g, device='cuda')
    got = op.to_grid_device(vd).cpu().numpy()
    for a, b in zip(got, v):
        assert rel_err(a, ref.WT.dot(b)) < 1e-12
    got = op.from_grid_device(gd).cpu().numpy()
    for a, b in zip(got, g):
        assert rel_err(a, ref.W.dot(b)) < 1e-12
    got = op.grid_mvm_device(gd).cpu().numpy()
    for a, b in zip(got, g):
        assert rel_err(a, ref.grid_matvec(b)) < 1e-11


def test_stated_size_minres_iterates_match_oracle(stated):
    """k = 1, 3, 7 iterations of the block solver on the full-width block against the oracle's restated
    scipy loop on the first pair and the odd last column."""
    import torch
    name, prob, op, ref, V = stated
    Vd = torch.as_tensor(V, device='cuda')
    cols = (0, 1, V.shape[0] - 1)
    for k in (1, 3, 7):
        X, iters, _, _ = op.minres_device(Vd, tol=1e-4, maxiter=k, check_every=10 ** 6)
        for c in cols:
            xr, _, itn_r, _ = orc.minres(ref.matvec, V[c], 1e-10, k)
            assert iters[c] == itn_r
            assert rel_err(X[c].cpu().numpy(), xr) < 1e-9


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


def test_bench_line_contract():
    """bench.py end to end on a small workload (every leg but the CPU baseline): one JSON line with the
    contract's keys, parity green, secondary workload rows present."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--workload', 'C', '--steps', '5',
                          '--warmup', '3', '--no-cpu', '--no-ill'], capture_output=True, text=True, timeout=900,
                         cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'parity', 'layouts', 'configs'):
        assert key in line, key
    assert line['value'] > 0 and line['gpu_launches'] > 0 and line['parity']['ok']
    assert line['layouts']['max_abs_diff_between_layouts'] == 0.0
    assert sorted(c['workload'][0] for c in line['configs']) == ['A', 'B', 'D']
    assert all(c['parity']['ok'] for c in line['configs'])
