"""CPU-only tests: host-side logic of the runlmc mirror (interpolation, error
behaviour), the C-ABI surface, and the multi-rank probe sharding with gloo."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest

from conftest import ROOT, load_golden
from runlmc_b200.approx import interpolation as itp


def test_interpolation_matches_reference_golden():
    g = load_golden('interp')
    for k in ('c1', 'c2', 'c3'):
        W = itp.interp_cubic(g[k + '_grid'], g[k + '_sample'])
        np.testing.assert_allclose(W.toarray(), g[k + '_dense'], rtol=0, atol=1e-15)
    W = itp.interp_bicubic(g['b1_gx'], g['b1_gy'], g['b1_sample'])
    np.testing.assert_allclose(W.toarray(), g['b1_dense'], rtol=0, atol=2e-15)
    W = itp.multi_interpolant([g['m1_sa'], g['m1_sb']], g['m1_grid'])
    np.testing.assert_allclose(W.toarray(), g['m1_dense'], rtol=0, atol=1e-15)
    assert W.lmc_geometry is not None
    W = itp.multi_interpolant([g['m2_sa'], g['m2_sb']], g['m2_gx'], g['m2_gy'])
    np.testing.assert_allclose(W.toarray(), g['m2_dense'], rtol=0, atol=2e-15)
    Xs = [np.arange(10.0).reshape(-1, 1), np.arange(4.0, 12).reshape(-1, 1)]
    np.testing.assert_allclose(itp.autogrid(Xs, None, None, None)[0], g['ag_a'])
    np.testing.assert_allclose(itp.autogrid(Xs, [-1], [13], [10])[0], g['ag_b'])


def test_interpolation_reference_unit_cases():
    # approx/test_interpolation.py:17-64 restated
    with pytest.raises(ValueError):
        itp.cubic_kernel(np.arange(0, 3, 0.1))
    with pytest.raises(ValueError):
        itp.cubic_kernel(np.arange(-3, 0, 0.1))
    assert itp.cubic_kernel(np.arange(-2, 2, 0.5).reshape(2, 2, 2)).shape == (2, 2, 2)
    assert itp.cubic_kernel(np.array([])).size == 0
    with pytest.raises(ValueError):
        itp.interp_cubic(np.arange(10).reshape(2, -1), np.array([3]))
    with pytest.raises(ValueError):
        itp.interp_cubic(np.arange(10), np.array([3]).reshape(-1, 1))
    with pytest.raises(ValueError):
        itp.interp_cubic(np.arange(3), np.array([1]))
    for m, n in [(10, 5), (10, 20), (4, 4)]:
        assert itp.interp_cubic(np.linspace(10, 20, m), np.logspace(10, 20, n)).shape == (n, m)
    assert itp.interp_cubic(np.arange(10.), np.array([])).shape == (0, 10)
    g = np.arange(10)
    with pytest.raises(ValueError):
        itp.interp_bicubic(g.reshape(2, -1), g, np.array([[3, 3]]))
    with pytest.raises(ValueError):
        itp.interp_bicubic(g, g, np.array([[3], [3]]))
    with pytest.raises(ValueError):
        itp.interp_bicubic(np.arange(3), g, np.array([[1, 1]]))
    # 2-D autogrid bounds (test_interpolation.py:207-229)
    x1 = np.column_stack([np.arange(10), np.arange(-3, 7)])
    x2 = np.column_stack([np.arange(4, 12), np.arange(2, 10)])
    for lo, hi, m in [([-1, -5], [13, 15], [10, 12]), ([5, 3], [9, 12], [3, 4]), (None, None, None)]:
        gx, gy = itp.autogrid([x1, x2], lo, hi, m)
        assert gx[1] <= 0 and gx[-2] >= 11 and gy[1] <= -3 and gy[-2] >= 10


def test_linalg_constructor_errors():
    # test_toeplitz.py:42-50, test_bttb.py:73-83, test_sum_matrix.py:91-98
    from runlmc_b200.linalg import Toeplitz, BTTB, SumMatrix, NumpyMatrix, Diag, SymmSquareBlockMatrix
    two_d = np.arange(8).reshape(2, 4)
    with pytest.raises(ValueError):
        Toeplitz(two_d)
    with pytest.raises(ValueError):
        Toeplitz(np.array([]))
    with pytest.raises(Exception):
        Toeplitz(np.arange(5) * 1j)
    with pytest.raises(ValueError):
        BTTB(two_d, two_d.shape)
    with pytest.raises(ValueError):
        BTTB(np.array([]), (0,))
    with pytest.raises(ValueError):
        BTTB(np.arange(8), (3, 4))
    with pytest.raises(Exception):
        BTTB(np.arange(5) * 1j, (5,))
    with pytest.raises(ValueError):
        SumMatrix([])
    with pytest.raises(ValueError):
        SumMatrix([NumpyMatrix(np.identity(i)) for i in [3, 3, 4]])
    with pytest.raises(ValueError):
        NumpyMatrix(np.arange(3))
    with pytest.raises(ValueError):
        Diag(np.ones((2, 2)))
    with pytest.raises(ValueError):
        SymmSquareBlockMatrix([[NumpyMatrix(np.identity(2))], []])
    t = Toeplitz(np.array([3.0, 1.0, 0.5]))
    assert t.shape == (3, 3) and t.dtype == np.float64
    np.testing.assert_array_equal(t.as_numpy(), [[3, 1, .5], [1, 3, 1], [.5, 1, 3]])
    assert t.upper_eig_bound() >= 5.0
    top = np.array([[4, 3, 2, 1], [3, 2, 1, 0], [2, 1, 0, 0]], dtype=float)
    b = BTTB(top.ravel(), top.shape).as_numpy()
    assert b.shape == (12, 12) and b[0, 5] == 2 and b[7, 2] == 2


def test_library_exports_every_declared_symbol():
    from runlmc_b200 import _native
    hdr = open(os.path.join(ROOT, 'include', 'lmc_b200.h')).read()
    names = set(re.findall(r'\b(lmc_[a-z0-9_]+)\s*\(', hdr))
    assert len(names) >= 25
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(names):
        assert hasattr(lib, name), name
    assert lib.lmc_version() >= 100
    # and the ctypes table covers the header
    assert names <= set(_native.SIGNATURES), names - set(_native.SIGNATURES)


def test_no_product_import_of_oracle():
    pkg = os.path.join(ROOT, 'runlmc_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('no CPU fallback', ''), os.path.join(dirpath, f)


def test_shard_bounds():
    from runlmc_b200.distributed import shard_bounds
    for N in (0, 1, 5, 16, 64, 127, 128):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_bounds(N, r, world)
                assert lo % 2 == 0 or lo == N
                cover.extend(range(lo, hi))
            assert cover == list(range(N))
    assert shard_bounds(128, 3, 8) == (48, 64)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from runlmc_b200.distributed import allreduce_trace, shard_bounds
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    per_probe = rng.standard_normal((10, 3, 2, 2))          # pretend per-probe Gram contributions
    lo, hi = shard_bounds(10, rank, world)
    trace = per_probe[lo:hi].sum(axis=0)
    ntrace = per_probe[lo:hi, 0, 0].sum(axis=0)
    out = allreduce_trace(trace, ntrace, float(hi - lo), 0.5 * (hi - lo))
    q.put((rank, out[0], out[1], out[2], out[3]))
    dist.destroy_process_group()


def test_allreduce_trace_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    rng = np.random.default_rng(0)
    per_probe = rng.standard_normal((10, 3, 2, 2))
    for rank, trace, ntrace, it, rs in res:
        np.testing.assert_allclose(trace, per_probe.sum(axis=0), rtol=1e-13)
        np.testing.assert_allclose(ntrace, per_probe[:, 0, 0].sum(axis=0), rtol=1e-13)
        assert it == 10.0 and rs == 5.0


def test_device_kernel_selection_rules():
    """gen_grid_kernel hands kernel evaluation to the device only for its own RBF / Matern32 /
    StdPeriodic classes on the operator's own grid distances (host logic, no GPU needed)."""
    from runlmc_b200 import kern
    from runlmc_b200.fused import kernel_descriptor, KERNEL_KINDS
    from runlmc_b200.lmc.grid_kernel import _device_kernels
    from runlmc_b200.lmc.functional_kernel import FunctionalKernel

    assert kernel_descriptor(kern.RBF(2.0)) == (KERNEL_KINDS['RBF'], 2.0, 1.0)
    assert kernel_descriptor(kern.Matern32(0.5)) == (KERNEL_KINDS['Matern32'], 0.5, 1.0)
    assert kernel_descriptor(kern.StdPeriodic(3.0, 0.25)) == (KERNEL_KINDS['StdPeriodic'], 3.0, 0.25)

    class MyRBF(kern.RBF):           # a subclass may override from_dist: evaluated on the host
        pass
    assert kernel_descriptor(MyRBF(1.0)) is None

    class FakeFused:
        def __init__(self, d):
            self._d = d

        def grid_dists(self):
            return self._d

    dists = np.linspace(0, 1, 33)
    fk = FunctionalKernel(D=2, lmc_kernels=[kern.RBF(1.0), kern.StdPeriodic(1.0, 0.5)], lmc_ranks=[1, 1])
    fk.set_input_dim(1)
    assert _device_kernels(fk, [0, 1], FakeFused(dists), dists) is not None
    assert _device_kernels(fk, [0, 1], FakeFused(dists), dists * (1 + 1e-9)) is None      # foreign distances
    assert _device_kernels(fk, [0, 1], FakeFused(dists), dists[:-1]) is None
    fk2 = FunctionalKernel(D=2, lmc_kernels=[MyRBF(1.0)], lmc_ranks=[1])
    fk2.set_input_dim(1)
    assert _device_kernels(fk2, [0], FakeFused(dists), dists) is None                     # foreign kernel class

    class Duck:                      # FunctionalKernel stand-ins without `_kernels` keep the uploaded tops
        pass
    assert _device_kernels(Duck(), [0], FakeFused(dists), dists) is None


def test_kernel_formulas_and_parameter_gradients():
    """runlmc_b200.kern (the formulas the device kernel evaluation restates, csrc/setup.cu): values
    at known points and parameter derivatives against central differences."""
    from runlmc_b200 import kern
    r = np.linspace(0.0, 2.5, 41)
    k = kern.RBF(3.0)
    np.testing.assert_allclose(k.from_dist(r), np.exp(-1.5 * r ** 2), rtol=1e-15)
    m = kern.Matern32(0.7)
    np.testing.assert_allclose(m.from_dist(r), (1 + np.sqrt(3) * 0.7 * r) * np.exp(-np.sqrt(3) * 0.7 * r), rtol=1e-14)
    p = kern.StdPeriodic(2.0, 0.8)
    np.testing.assert_allclose(p.from_dist(r), np.exp(-np.sin(np.pi * r / 0.8) ** 2), rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(p.from_dist(r + 0.8), p.from_dist(r), rtol=1e-9, atol=1e-12)     # period
    assert k.from_dist(0.0) == m.from_dist(0.0) == p.from_dist(0.0) == 1.0

    def fd(make, params, i, h=1e-6):
        lo, hi = list(params), list(params)
        lo[i] -= h
        hi[i] += h
        return (make(*hi).from_dist(r) - make(*lo).from_dist(r)) / (2 * h)

    for make, params in ((kern.RBF, [3.0]), (kern.Matern32, [0.7]), (kern.StdPeriodic, [2.0, 0.8])):
        grads = make(*params).kernel_gradient(r)
        assert len(grads) == len(params) == len(make(*params).param_values())
        for i, g in enumerate(grads):
            np.testing.assert_allclose(g, fd(make, params, i), rtol=2e-6, atol=2e-8)


def test_logdet_quadrature_host_side():
    """The host half of the stochastic log-determinant (eigen-decomposition of the recorded tridiagonals) against
    the oracle's quadrature on coefficients recorded by the oracle's MINRES loop, including a column that stopped
    early and a record shorter than the solve."""
    from oracle import lmc_oracle as orc
    from runlmc_b200.approx.logdet import quadrature_terms
    rng = np.random.default_rng(4)
    n = 60
    A = rng.standard_normal((n, n))
    A = A.dot(A.T) + n * np.eye(n)
    recs, want = [], []
    for k in (9, 5):
        z = rng.integers(0, 2, n) * 2.0 - 1.0
        rec = []
        orc.minres(A.dot, z, 1e-30, k, record=rec)
        recs.append(rec)
        want.append(orc.lanczos_quadrature_logdet(rec[0], rec[1:]))
    K = 12
    tri = np.zeros((2, K, 2))
    for c, rec in enumerate(recs):
        tri[c, :len(rec) - 1] = np.array(rec[1:])
    beta1 = np.array([r[0] for r in recs])
    iters = np.array([len(r) - 1 for r in recs])
    got = quadrature_terms(tri, beta1, iters)
    assert np.allclose(got, want, rtol=1e-12)
    # a record of 4 steps of the 9-step solve uses the leading 4 x 4 tridiagonal
    short = quadrature_terms(tri[:1, :4], beta1[:1], iters[:1])
    assert np.isclose(short[0], orc.lanczos_quadrature_logdet(recs[0][0], recs[0][1:5]), rtol=1e-12)
    # a zero right-hand side (no iterations) contributes nothing
    assert quadrature_terms(tri[:1], beta1[:1], np.array([0]))[0] == 0.0
    # Gauss quadrature with n nodes of a log-like integrand: sanity against the exact value for this small matrix
    exact = np.linalg.slogdet(A)[1]
    probes = rng.integers(0, 2, (200, n)) * 2.0 - 1.0
    est, _ = orc.stochastic_logdet(A.dot, probes, 1e-12, n)
    assert abs(est - exact) < 0.02 * abs(exact)
