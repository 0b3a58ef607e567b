"""GPU parity tests of the fused CUDA operator against the CPU oracle and the
reference golden vectors.  All calls go through the C ABI (ctypes)."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import lmc_oracle as orc
from runlmc_b200 import synthetic

pytestmark = pytest.mark.gpu

MVM_TOL = 1e-10      # north_star: MVMs within 1e-10 relative
GRAD_TOL = 1e-8      # north_star: gradient within 1e-8 relative (identical probes)


def fused_from_problem(prob, factors=True):
    """factors=True also hands over the LMC factors B_q = A_q^T A_q + diag(kappa_q),
    which selects the low-rank spectral mix where it is cheaper."""
    from runlmc_b200.fused import FusedLMC
    op = FusedLMC(prob.Xs, prob.grids)
    if factors:
        op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    else:
        op.set_params(prob.tops, prob.coreg_mats(), prob.noise)
    return op


def oracle_from_problem(prob, rep='sum'):
    spec = orc.KernelSpec(['rbf'] * prob.Q, [[g] for g in prob.gammas],
                          prob.coreg_vecs, prob.coreg_diags, prob.noise)
    return spec, orc.build_operator(spec, prob.Xs, prob.grids, rep=rep)


PROBLEMS = {
    'A_edge': lambda: synthetic.make_problem('A', seed=1234, edge=True, cells_per_lengthscale=4),
    'A': lambda: synthetic.make_problem('A', seed=5, cells_per_lengthscale=4),
    '2d_small': lambda: synthetic.make_problem('e_small', seed=99, cells_per_lengthscale=3,
                                               lens=[120, 90, 100], grid=[12, 10], N=5),
    '2d_edge': lambda: synthetic.make_problem('e_small', seed=7, cells_per_lengthscale=3, edge=True),
    '2d_rect': lambda: synthetic.make_problem('e_small', seed=8, cells_per_lengthscale=3,
                                              lens=[500, 0, 433], grid=[40, 9], N=3),
    'd_small': lambda: synthetic.make_problem('d_small', seed=11, cells_per_lengthscale=6),
    'ragged': lambda: synthetic.make_problem('d_small', seed=12, cells_per_lengthscale=6,
                                             lens=[1, 350, 0, 77], grid=[100], N=3),
    'C': lambda: synthetic.make_problem('C', seed=13, cells_per_lengthscale=5),
    # 128 < m_x <= 256: the register/shuffle 512-point column kernel (spectral_col512.cuh), m_x not a
    # multiple of 32 so the last loads/stores of a line are partially masked
    'col512': lambda: synthetic.make_problem('e_small', seed=15, cells_per_lengthscale=4,
                                             lens=[900, 800, 850], grid=[200, 12], N=6),
    # both axes in (128, 256]: register transforms + bulk-copy (TMA) row passes (spectral_rows512.cuh); m_x not
    # a multiple of 8 (partial last row group), m_y not a multiple of 32
    'rows512': lambda: synthetic.make_problem('e_small', seed=17, cells_per_lengthscale=4, edge=True,
                                              lens=[700, 0, 650], grid=[130, 200], N=6),
    'col512_edge': lambda: synthetic.make_problem('e_small', seed=16, cells_per_lengthscale=4, edge=True,
                                                  lens=[400, 0, 300, 500, 20], D=5, grid=[131, 9], N=4),
    'four_step': lambda: synthetic.make_problem('d_small', seed=14, cells_per_lengthscale=40,
                                                lens=[3000, 2500], D=2, grid=[5000], Q=2, N=3),
}


@pytest.mark.parametrize('name', sorted(PROBLEMS))
def test_stages_and_mvm(name):
    import torch
    prob = PROBLEMS[name]()
    op = fused_from_problem(prob)
    _, ref = oracle_from_problem(prob)
    rng = np.random.default_rng(0)
    P = 5
    V = rng.standard_normal((P, prob.n))
    G = rng.standard_normal((P, prob.D * ref.m))
    Vd = torch.as_tensor(V, device='cuda')
    Gd = torch.as_tensor(G, device='cuda')
    got = op.to_grid_device(Vd).cpu().numpy()
    want = np.array([ref.WT.dot(v) for v in V])
    assert rel_err(got, want) < 1e-13
    got = op.from_grid_device(Gd).cpu().numpy()
    want = np.array([ref.W.dot(g) for g in G])
    assert rel_err(got, want) < 1e-13
    got = op.grid_mvm_device(Gd).cpu().numpy()
    want = np.array([ref.grid_matvec(g) for g in G])
    assert rel_err(got, want) < 1e-12
    got = op.mvm(V)
    want = np.array([ref.matvec(v) for v in V])
    for g, w in zip(got, want):
        assert rel_err(g, w) < MVM_TOL
    # single vector / matmat entry points
    assert rel_err(op.matvec(V[0]), want[0]) < MVM_TOL
    assert rel_err(op.matmat(V.T), want.T) < MVM_TOL
    # device entry point, odd and even block widths
    for p in (1, 2, 4):
        out = op.mvm_device(Vd[:p].contiguous()).cpu().numpy()
        assert rel_err(out, want[:p]) < MVM_TOL
    # the solver's layout: vectors in the operator's sorted point order
    perm = op.perm()
    assert sorted(perm.tolist()) == list(range(prob.n))
    Vs = torch.as_tensor(np.ascontiguousarray(V[:, perm]), device='cuda')
    out = op.mvm_sorted_device(Vs).cpu().numpy()
    assert rel_err(out, want[:, perm]) < MVM_TOL


@pytest.mark.parametrize('name', sorted(PROBLEMS))
@pytest.mark.parametrize('P', [1, 4, 7])
def test_point_major_blocks(name, P):
    """lmc_mvm_rows (X[i][c], the reference's matmat argument as numpy lays it out) against the oracle and,
    bit for bit, against the column-major entry point -- on every geometry, i.e. through the direct row
    staging of the 2-D strip scatter and through the transposing fallback of all other scatters."""
    import torch
    prob = PROBLEMS[name]()
    op = fused_from_problem(prob)
    _, ref = oracle_from_problem(prob)
    rng = np.random.default_rng(5)
    X = rng.standard_normal((prob.n, P))
    Xd = torch.as_tensor(X, device='cuda')
    Y = op.matmat_device(Xd).cpu().numpy()
    for c in range(P):
        assert rel_err(Y[:, c], ref.matvec(X[:, c])) < MVM_TOL
    assert np.array_equal(Y, op.mvm_device(Xd.t().contiguous()).cpu().numpy().T)
    assert np.array_equal(op.matmat(X), Y)
    # a column slice of a wider block: row stride > P, rows that start 8 bytes off a 16-byte boundary
    wide_in = torch.full((prob.n, P + 6), float('nan'), dtype=torch.float64, device='cuda')
    wide_out = torch.full((prob.n, P + 4), float('nan'), dtype=torch.float64, device='cuda')
    wide_in[:, 3:3 + P] = Xd
    op.matmat_device(wide_in[:, 3:3 + P], wide_out[:, 1:1 + P])
    assert np.array_equal(wide_out[:, 1:1 + P].cpu().numpy(), Y)
    assert torch.isnan(wide_out[:, 0]).all() and torch.isnan(wide_out[:, 1 + P:]).all()     # nothing else written


@pytest.mark.parametrize('name', ['lmc_A', 'lmc_2d', 'lmc_B'])
def test_mvm_against_reference_golden(name):
    from test_oracle_golden import GOLDEN_PROBLEMS
    g = load_golden(name)
    prob = GOLDEN_PROBLEMS[name]()
    op = fused_from_problem(prob)
    got = op.mvm(g['V'])
    for a, b in zip(got, g['KV']):
        assert rel_err(a, b) < MVM_TOL


@pytest.mark.parametrize('name', ['2d_small', 'd_small', 'C', 'four_step', 'col512', 'col512_edge', 'rows512'])
def test_lowrank_mix_equals_dense_mix(name):
    prob = PROBLEMS[name]()
    rng = np.random.default_rng(3)
    V = rng.standard_normal((3, prob.n))
    a = fused_from_problem(prob, factors=True).mvm(V)
    b = fused_from_problem(prob, factors=False).mvm(V)
    assert rel_err(a, b) < 1e-13
    from runlmc_b200.fused import FusedLMC
    op = FusedLMC(prob.Xs, prob.grids)
    with pytest.raises(ValueError):        # factors must reproduce B
        op.set_params(prob.tops, prob.coreg_mats(), prob.noise,
                      [2 * a_ for a_ in prob.coreg_vecs], prob.coreg_diags)


def test_linearity_and_symmetry_large():
    """Size-independent properties at a size the oracle would take long on."""
    import torch
    prob = synthetic.make_problem('E', seed=3, cells_per_lengthscale=6,
                                  lens=[20000] * 4, D=4, grid=[128, 64], N=2)
    op = fused_from_problem(prob)
    rng = np.random.default_rng(1)
    V = torch.as_tensor(rng.standard_normal((4, prob.n)), device='cuda')
    KV = op.mvm_device(V)
    a, b = 0.7, -1.3
    lin = op.mvm_device((a * V[0] + b * V[1]).reshape(1, -1).contiguous())[0]
    assert rel_err((a * KV[0] + b * KV[1]).cpu().numpy(), lin.cpu().numpy()) < 1e-12
    # symmetry: u' K v == v' K u
    s1 = float(torch.dot(V[2], KV[3])); s2 = float(torch.dot(V[3], KV[2]))
    assert abs(s1 - s2) <= 1e-11 * max(abs(s1), 1.0)
    # positive definiteness on these vectors
    assert float(torch.dot(V[0], KV[0])) > 0


@pytest.mark.parametrize('ndim', [2, 1])
def test_wide_blocks_match_narrow_blocks_and_oracle(ndim):
    """Block widths that exercise what the 4-column tests cannot: several 16-pair groups per CTA, the
    odd pair(s) routed through the one-pair scatter kernel (33 = 2*16 + 1 pairs, 34 = 2*16 + 2), several
    8-pair gather passes, more than one block of 128 pairs, the pipelined row pass (many slabs)."""
    import torch
    if ndim == 2:
        # ~1.3 points per bin: the tiled scatter / gather kernels' capacities hold, like at config E
        prob = synthetic.make_problem('E', seed=4, cells_per_lengthscale=5, lens=[4000, 3500, 3800],
                                      D=3, grid=[64, 48], N=2)
    else:
        prob = synthetic.make_problem('D', seed=4, cells_per_lengthscale=5, lens=[9000, 7000, 8000],
                                      D=3, grid=[1024], N=2)
    op = fused_from_problem(prob)
    _, ref = oracle_from_problem(prob)
    rng = np.random.default_rng(2)
    Pmax = 300
    V = rng.standard_normal((Pmax, prob.n))
    Vd = torch.as_tensor(V, device='cuda')
    narrow = torch.cat([op.mvm_device(Vd[i:i + 4].contiguous()) for i in range(0, Pmax, 4)]).cpu().numpy()
    for c in (0, 65, 66, 128, 299):
        assert rel_err(narrow[c], ref.matvec(V[c])) < MVM_TOL
    perm = op.perm()
    # 8, 9, 10, 17, 20, 25, 33, 34, 65, 150 pairs; 17 / 33 / 49 / 65 / 129 columns are 8 k pairs + one odd
    # column, which the 2-D scatter carries as a third accumulator of the last group
    for P in (16, 17, 19, 33, 40, 49, 65, 67, 129, 300):
        wide = op.mvm_device(Vd[:P].contiguous()).cpu().numpy()
        assert rel_err(wide, narrow[:P]) < 1e-12
        assert max(rel_err(wide[c], narrow[c]) for c in range(P)) < 1e-12
        Vs = torch.as_tensor(np.ascontiguousarray(V[:P][:, perm]), device='cuda')
        wide_sorted = op.mvm_sorted_device(Vs).cpu().numpy()
        assert max(rel_err(wide_sorted[c], narrow[c][perm]) for c in range(P)) < 1e-12
        # point-major block ([n, P], numpy's C order): same staged values, same summation order
        rows = op.matmat_device(Vd[:P].t().contiguous()).cpu().numpy()
        assert np.array_equal(rows, wide.T)
    # host-buffer entry points on a wide block
    assert rel_err(op.mvm(V[:67]), narrow[:67]) < 1e-12
    X = np.ascontiguousarray(V[:67].T)
    assert rel_err(op.matmat(X), op.mvm(V[:67]).T) < 1e-12     # the column path cuts the block into chunks
    # the host entry point pipelines chunks of 32 columns: other pair groups, same numbers to rounding
    assert rel_err(op.matmat(X), op.matmat_device(torch.as_tensor(X, device='cuda')).cpu().numpy()) < 1e-12
    assert np.array_equal(op.matmat(X[:, :33].copy()), op.matmat_device(torch.as_tensor(X[:, :33].copy(), device='cuda')).cpu().numpy())
    assert rel_err(op.matmat(np.asfortranarray(X)), op.matmat(X)) < 1e-12
    assert rel_err(op.matmat(X[:, :5]), op.matmat(X)[:, :5]) < 1e-12     # a strided view is copied first


@pytest.mark.parametrize('P', [35, 33])
def test_wide_minres_matches_single_column_solves(P):
    """Columns of a wide block stop at different iterations (activity masks inside the kernels);
    every column must end where it ends when solved alone."""
    prob = well_conditioned_problem()
    op = fused_from_problem(prob)
    rng = np.random.default_rng(3)
    RHS = rng.standard_normal((P, prob.n)) * np.logspace(-3, 1, P)[:, None]
    RHS[7] = 0.0
    X, iters, resid, istop = op.minres(RHS, tol=1e-4, check_every=5)
    assert len(set(iters.tolist())) > 2                     # they really stop at different times
    for c in (0, 7, 12, P - 2, P - 1):
        x1, it1, r1, st1 = op.minres(RHS[c:c + 1], tol=1e-4, check_every=5)
        assert it1[0] == iters[c] and st1[0] == istop[c]
        # a column shares its complex FFT with its pair partner: ulp-level cross-talk through the
        # twiddle products, amplified by Lanczos to ~1e-6 at convergence -- far inside the tolerance
        assert rel_err(X[c], x1[0]) < SOLVE_TOL or np.all(x1[0] == 0)
    assert np.all(resid < 1e-4)


@pytest.mark.parametrize('P', [37, 33, 17])
def test_wide_minres_2d(P):
    """P = 33 / 17: 16 / 8 pairs and an odd column that the scatter carries as a third accumulator, with
    columns (the odd one included) stopping at different iterations."""
    prob = PROBLEMS['2d_small']()
    prob.noise = np.full(prob.D, 25.0)
    op = fused_from_problem(prob)
    _, ref = oracle_from_problem(prob)
    rng = np.random.default_rng(5)
    RHS = rng.standard_normal((P, prob.n)) * np.logspace(-2, 1, P)[:, None]
    X, iters, resid, istop = op.minres(RHS, tol=1e-4, check_every=5)
    for c in (0, P // 2, P - 2, P - 1):
        xr, ctr, err = orc.iterative_solve(ref.matvec, RHS[c], 1e-4, check_every=5)
        assert abs(int(iters[c]) - ctr) <= 5
        assert rel_err(X[c], xr) < SOLVE_TOL
    assert np.all(resid < 1e-4)


# MINRES parity.  Lanczos amplifies ulp-level differences in the operator: the
# reference run with two of its own equivalent representations ('sum' vs 'bt')
# already differs by 1e-3..1e-1 in mid-convergence iterates (k ~ 25) on these
# problems, so iterates are compared (a) tightly at small fixed iteration
# counts, (b) tightly at larger counts on a well-conditioned problem, and (c)
# at convergence within the solver tolerance.
SOLVE_TOL = 1e-5     # relative agreement of converged solutions (solver tol = 1e-4 absolute residual)


@pytest.mark.parametrize('name', ['lmc_A', 'lmc_2d', 'lmc_B'])
def test_minres_against_reference_golden(name):
    from test_oracle_golden import GOLDEN_PROBLEMS
    g = load_golden(name)
    prob = GOLDEN_PROBLEMS[name]()
    op = fused_from_problem(prob)
    _, ref = oracle_from_problem(prob)
    RHS = np.vstack([prob.y[None, :], g['probes']])
    X, iters, resid, istop = op.minres(RHS, tol=1e-4)
    want = np.vstack([g['alpha'][None, :], g['inv_probes']])
    ref_it = int(g['solve_y_ctr'])
    assert abs(int(iters[0]) - ref_it) <= max(3, 0.03 * ref_it)
    for b, x, w, r in zip(RHS, X, want, resid):
        assert rel_err(x, w) < SOLVE_TOL
        ref_res = np.linalg.norm(b - ref.matvec(w))
        # reported residual is the true residual, and as good as the reference's
        assert abs(r - np.linalg.norm(b - ref.matvec(x))) <= 1e-9 + 1e-6 * r
        assert r <= max(3e-4, 5 * ref_res)


@pytest.mark.parametrize('name', ['A', '2d_small', 'd_small'])
def test_minres_fixed_iterations(name):
    """Equal iteration counts: compare iterates with the oracle's restated scipy loop."""
    prob = PROBLEMS[name]()
    op = fused_from_problem(prob)
    _, ref = oracle_from_problem(prob)
    RHS = np.vstack([prob.y[None, :], prob.probes[:3]])
    for k in (1, 2, 3, 7):
        X, iters, _, istop = op.minres(RHS, tol=1e-4, maxiter=k, check_every=10 ** 6)
        for b, x, it in zip(RHS, X, iters):
            xr, istop_r, itn_r, _ = orc.minres(ref.matvec, b, 1e-10, k)
            assert it == itn_r
            assert rel_err(x, xr) < 1e-10


def well_conditioned_problem():
    prob = synthetic.make_problem('d_small', seed=21, cells_per_lengthscale=6)
    prob.noise = np.full(prob.D, 40.0)      # cond(K~) of a few tens
    return prob


def test_minres_fixed_iterations_well_conditioned():
    """Later iterates: the yardstick is the reference's own spread between two
    of its equivalent operator representations ('sum' vs 'bt') at the same k."""
    prob = well_conditioned_problem()
    op = fused_from_problem(prob)
    _, ref = oracle_from_problem(prob)
    _, ref_bt = oracle_from_problem(prob, rep='bt')
    RHS = np.vstack([prob.y[None, :], prob.probes[:2]])
    for k in (10, 25):
        X, iters, _, istop = op.minres(RHS, tol=1e-4, maxiter=k, check_every=10 ** 6)
        for b, x, it, st in zip(RHS, X, iters, istop):
            xr, istop_r, itn_r, _ = orc.minres(ref.matvec, b, 1e-10, k)
            xb, _, _, _ = orc.minres(ref_bt.matvec, b, 1e-10, k)
            assert it == itn_r and st == istop_r
            assert rel_err(x, xr) < 1e-10 + 10 * rel_err(xb, xr)


def test_minres_zero_rhs_and_mixed():
    prob = PROBLEMS['A']()
    op = fused_from_problem(prob)
    RHS = np.vstack([np.zeros(prob.n), prob.y, 1e-3 * prob.probes[0]])
    X, iters, resid, istop = op.minres(RHS, tol=1e-4)
    # scipy returns x = 0 immediately for a zero right-hand side
    assert iters[0] == 0 and np.all(X[0] == 0) and resid[0] < 1e-10
    _, ref = oracle_from_problem(prob)
    for b, x, it in zip(RHS[1:], X[1:], iters[1:]):
        xr, ctr, err = orc.iterative_solve(ref.matvec, b, 1e-4)
        assert abs(int(it) - ctr) <= max(3, 0.03 * ctr)
        # tol is an absolute residual: loose relative to the 1e-3-scaled rhs
        assert rel_err(x, xr) < 10 * SOLVE_TOL


def test_minres_residual_check_terminates():
    """The reference's every-100-iterations true-residual test (iterative.py:36-42),
    exercised with a short period."""
    prob = well_conditioned_problem()
    op = fused_from_problem(prob)
    _, ref = oracle_from_problem(prob)
    RHS = np.vstack([prob.y[None, :], prob.probes[:3]])
    X, iters, resid, istop = op.minres(RHS, tol=1e-4, check_every=5)
    for b, x, it, r, st in zip(RHS, X, iters, resid, istop):
        xr, ctr, err = orc.iterative_solve(ref.matvec, b, 1e-4, check_every=5)
        assert abs(int(it) - ctr) <= 5 and it % 5 == 0 and st == 10
        assert r < 1e-4 and rel_err(x, xr) < SOLVE_TOL


@pytest.mark.parametrize('name', ['lmc_A', 'lmc_2d', 'lmc_B'])
def test_gradient_against_reference_golden(name):
    from test_oracle_golden import GOLDEN_PROBLEMS
    from runlmc_b200.fused import assemble_gradients
    g = load_golden(name)
    prob = GOLDEN_PROBLEMS[name]()
    op = fused_from_problem(prob)
    extra = [t for ts in prob.top_grads for t in ts]
    quad, trace, nquad, ntrace = op.grad_grams(g['alpha'], g['probes'], g['inv_probes'], extra)
    cv, cd, kg, nz = assemble_gradients(prob.coreg_vecs, prob.coreg_mats(),
                                        [len(t) for t in prob.top_grads], prob.N,
                                        quad, trace, nquad, ntrace)
    assert rel_err(np.array(cv), g['g_coreg_vec']) < GRAD_TOL
    assert rel_err(np.array(cd), g['g_coreg_diag']) < GRAD_TOL
    assert rel_err(np.array(kg), g['g_kernel']) < GRAD_TOL
    assert rel_err(nz, g['g_noise']) < GRAD_TOL


@pytest.mark.parametrize('name', ['lmc_B', 'lmc_2d'])
def test_sharded_gradient_single_rank(name):
    """The whole gradient evaluation (device-resident solves + Gram stage + chain rule) through
    distributed.sharded_gradient with one rank; the golden solves stopped at the same iterates the
    reference's did, so gradients agree to the solver tolerance, not to GRAD_TOL."""
    from test_oracle_golden import GOLDEN_PROBLEMS
    from runlmc_b200.distributed import sharded_gradient
    g = load_golden(name)
    prob = GOLDEN_PROBLEMS[name]()
    op = fused_from_problem(prob)
    grads, stats = sharded_gradient(op, prob.y, g['probes'], prob.top_grads, prob.coreg_vecs,
                                    prob.coreg_mats(), tol=1e-4)
    assert rel_err(stats['alpha'], g['alpha']) < 10 * SOLVE_TOL
    assert rel_err(np.array(grads[0]), g['g_coreg_vec']) < 1e-3
    assert rel_err(np.array(grads[1]), g['g_coreg_diag']) < 1e-3
    assert rel_err(np.array(grads[2]), g['g_kernel']) < 1e-3
    assert rel_err(grads[3], g['g_noise']) < 1e-3


def test_error_paths():
    from runlmc_b200.fused import FusedLMC
    prob = PROBLEMS['A']()
    with pytest.raises(ValueError):
        FusedLMC(prob.Xs, [np.linspace(0, 1, 3)])          # grid < 4 (interpolation.py:91-92)
    with pytest.raises(ValueError):
        FusedLMC(prob.Xs, [np.linspace(0, 1, 10).reshape(2, 5)])
    op = FusedLMC(prob.Xs, prob.grids)
    with pytest.raises(ValueError):
        op.mvm(np.zeros(prob.n))                            # parameters not set
    with pytest.raises(ValueError):
        op.set_params(prob.tops, prob.coreg_mats()[:1], prob.noise)
    op.set_params(prob.tops, prob.coreg_mats(), prob.noise)
    with pytest.raises(ValueError):
        op.mvm(np.zeros(prob.n + 1))
    import torch
    empty = torch.empty((0, prob.n), dtype=torch.float64, device='cuda')
    assert op.mvm_device(empty).shape == (0, prob.n)          # empty blocks are valid products
    assert op.mvm_sorted_device(empty).shape == (0, prob.n)
    assert op.mvm(np.zeros((0, prob.n))).shape == (0, prob.n)


@pytest.mark.parametrize('name', ['lmc_A', 'lmc_2d', 'lmc_B'])
def test_cg_against_reference_golden(name):
    """Iterative.solve(..., minres=False): the device CG against the reference's own run
    (tests/golden/cg.npz) and, column by column in a block, against the oracle's restated scipy loop."""
    from test_oracle_golden import GOLDEN_PROBLEMS
    g = load_golden('cg')
    prob = GOLDEN_PROBLEMS[name]()
    op = fused_from_problem(prob)
    _, ref = oracle_from_problem(prob)
    RHS = np.vstack([prob.y[None, :], prob.probes[:3], np.zeros((1, prob.n))])
    X, iters, resid, info = op.cg(RHS, tol=1e-4)
    # iteration counts: CG's own stop ||r|| < 1e-10 ||b|| sits on the recurrence residual, which
    # rounding moves by a few iterations on the ill-conditioned case (lmc_B: 40 in the reference)
    slack = lambda ctr: max(1, ctr // 8)   # noqa: E731
    assert abs(int(iters[0]) - int(g[name + '_ctr'])) <= slack(int(g[name + '_ctr']))
    assert rel_err(X[0], g[name + '_x']) < SOLVE_TOL
    assert resid[0] <= max(1e-4, 3 * float(g[name + '_err']))
    for b, x, it, r in zip(RHS[1:4], X[1:4], iters[1:4], resid[1:4]):
        xr, ctr, err = orc.iterative_solve(ref.matvec, b, 1e-4, use_minres=False)
        assert abs(int(it) - ctr) <= slack(ctr)
        assert rel_err(x, xr) < SOLVE_TOL and r <= max(1e-4, 3 * err)
        assert abs(np.linalg.norm(b - ref.matvec(x)) - r) <= 1e-8 * max(1.0, np.linalg.norm(b))
    assert iters[4] == 0 and not X[4].any() and info[4] == 0          # zero right-hand side


def test_cg_fixed_iterations_match_oracle():
    """Iterates of the first CG steps agree with the restated scipy loop to rounding."""
    prob = PROBLEMS['d_small']()
    op = fused_from_problem(prob)
    _, ref = oracle_from_problem(prob)
    b = prob.probes[0]
    for k in (1, 2, 5):
        X, iters, _, info = op.cg(b, tol=1e-4, maxiter=k)
        xr, _, _ = orc.cg(ref.matvec, b, 1e-10, k)
        assert iters[0] == k and info[0] == k
        assert rel_err(X[0], xr) < 1e-10


def test_two_handles_from_two_threads():
    """Handles own their solver state: two operators solving at the same time from two host
    threads (ctypes releases the GIL) give the results of the same solves run one after the other."""
    import threading
    pa, pb = PROBLEMS['d_small'](), PROBLEMS['2d_edge']()
    oa, ob = fused_from_problem(pa), fused_from_problem(pb)
    Ra = np.vstack([pa.y[None, :], pa.probes[:4]])
    Rb = np.vstack([pb.y[None, :], pb.probes[:3]])
    want_a, want_b = oa.minres(Ra, tol=1e-4), ob.cg(Rb, tol=1e-4)
    got = {}
    for rep in range(3):
        ta = threading.Thread(target=lambda: got.__setitem__('a', oa.minres(Ra, tol=1e-4)))
        tb = threading.Thread(target=lambda: got.__setitem__('b', ob.cg(Rb, tol=1e-4)))
        ta.start(); tb.start(); ta.join(); tb.join()
        for k, want in (('a', want_a), ('b', want_b)):
            np.testing.assert_array_equal(got[k][0], want[0])     # deterministic kernels, private state
            np.testing.assert_array_equal(got[k][1], want[1])


@pytest.mark.parametrize('base,Q,ranks', [
    ('d_small', 1, [1]), ('d_small', 5, [1, 2, 1, 2, 1]), ('d_small', 7, [1] * 7), ('d_small', 3, [3, 1, 2]),
    ('e_small', 1, [2]), ('e_small', 5, [2, 1, 1, 2, 1]), ('e_small', 6, [1, 1, 3, 1, 1, 1]),
])
def test_many_kernels_and_higher_ranks(base, Q, ranks):
    """Q = 1, Q > 4 (run-time kernel count in the fused mix) and coregionalisation ranks 1..3 (rank 3
    takes the dense mix) against the oracle, with and without the low-rank factors; products,
    a few solver iterations and the gradient Gram stage."""
    from runlmc_b200.fused import FusedLMC, assemble_gradients
    prob = synthetic.make_problem(base, seed=31, cells_per_lengthscale=4, Q=Q)
    rng = np.random.default_rng(7)
    prob.coreg_vecs = [rng.uniform(-1, 1, size=(r, prob.D)) for r in ranks]
    spec, ref = oracle_from_problem(prob)
    V = rng.standard_normal((5, prob.n))
    want = np.array([ref.matvec(v) for v in V])
    ops = []
    for factors in (True, False):
        op = fused_from_problem(prob, factors=factors)
        assert rel_err(op.mvm(V), want) < MVM_TOL
        ops.append(op)
    X, iters, _, _ = ops[0].minres(V[:2], tol=1e-4, maxiter=4)
    for b, x in zip(V[:2], X):
        xr, _, _, _ = orc.minres(ref.matvec, b, 1e-10, 4)
        assert rel_err(x, xr) < 1e-9
    # Gram stage: identical inputs, low-rank vs dense operator and the oracle's gradient of the same
    alpha, R, RINV = V[0], prob.probes[:4], V[1:5]
    extra = [t for ts in prob.top_grads for t in ts]
    g0 = ops[0].grad_grams(alpha, R, RINV, extra)
    g1 = ops[1].grad_grams(alpha, R, RINV, extra)
    for a, b in zip(g0, g1):
        assert rel_err(a, b) < 1e-11


@pytest.mark.parametrize('ndim', [1, 2])
def test_dense_and_clustered_points_against_oracle(ndim):
    """Point sets far denser than the tiled kernels' shared-memory capacities, with thousands of
    points inside single cells and at the grid edge: the 1-D scatter streams a tile in several
    chunks, the 2-D scatter/gather fall back to their capacity-free kernels.  Products, both
    interpolation stages and the first solver iterates against the oracle."""
    import torch
    rng = np.random.default_rng(17)
    if ndim == 1:
        prob = synthetic.make_problem('d_small', seed=5, cells_per_lengthscale=6,
                                      lens=[4000, 3000, 10, 2500], grid=[64], N=3)
        prob.Xs[0][:1500] = 0.5 + 0.01 * rng.uniform(size=(1500, 1))          # inside one cell
        prob.Xs[3][:800] = 1.0 - 1e-3 * rng.uniform(size=(800, 1))            # clamped stencils at the edge
    else:
        prob = synthetic.make_problem('e_small', seed=6, cells_per_lengthscale=3,
                                      lens=[3000, 2500, 2800], grid=[12, 10], N=3)
        prob.Xs[0][:2000] = np.array([0.41, 0.63]) + 0.02 * rng.uniform(size=(2000, 2))
        prob.Xs[2][:600] = np.array([0.999, 0.001]) + 1e-3 * rng.uniform(-1, 0, size=(600, 2)) * [1, -1]
    _, ref = oracle_from_problem(prob)
    V = rng.standard_normal((5, prob.n))
    G = rng.standard_normal((3, prob.D * ref.m))
    want = np.array([ref.matvec(v) for v in V])
    from runlmc_b200.fused import FusedLMC
    for build in ('device', 'host'):
        op = FusedLMC(prob.Xs, prob.grids, build=build)
        op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
        assert rel_err(op.mvm(V), want) < MVM_TOL
        Vd, Gd = torch.as_tensor(V, device='cuda'), torch.as_tensor(G, device='cuda')
        assert rel_err(op.to_grid_device(Vd).cpu().numpy(), np.array([ref.WT.dot(v) for v in V])) < 1e-12
        assert rel_err(op.from_grid_device(Gd).cpu().numpy(), np.array([ref.W.dot(g) for g in G])) < 1e-12
    X, iters, _, _ = op.minres(V[:2], tol=1e-4, maxiter=5)
    for b, x in zip(V[:2], X):
        xr, _, _, _ = orc.minres(ref.matvec, b, 1e-10, 5)
        assert rel_err(x, xr) < 1e-9


@pytest.mark.parametrize('ndim', [1, 2])
def test_point_major_blocks_identity_permutation(ndim):
    """Inputs that already are in the operator's point order (sorted by output and grid bin): the permutation is
    the identity, the kernels index without it, and the point-major entry point takes its transposing fall-back
    for the scatter."""
    import torch
    if ndim == 1:
        prob = synthetic.make_problem('d_small', seed=21, cells_per_lengthscale=6)
        for X in prob.Xs:
            X[:] = np.sort(X, axis=0)
    else:
        prob = synthetic.make_problem('e_small', seed=21, cells_per_lengthscale=3)
        m0, m1 = (len(g) for g in prob.grids)
        for X in prob.Xs:            # sort by (bin of axis 0, bin of axis 1), the operator's own key
            fx = np.floor(X[:, 0] * (m0 - 1)).astype(int)
            fy = np.floor(X[:, 1] * (m1 - 1)).astype(int)
            X[:] = X[np.lexsort((fy, fx))]
    op = fused_from_problem(prob)
    assert np.array_equal(op.perm(), np.arange(prob.n))
    _, ref = oracle_from_problem(prob)
    rng = np.random.default_rng(6)
    X = rng.standard_normal((prob.n, 5))
    Y = op.matmat_device(torch.as_tensor(X, device='cuda')).cpu().numpy()
    for c in range(5):
        assert rel_err(Y[:, c], ref.matvec(X[:, c])) < MVM_TOL
    assert np.array_equal(Y, op.mvm_device(torch.as_tensor(np.ascontiguousarray(X.T), device='cuda')).cpu().numpy().T)
