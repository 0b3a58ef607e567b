"""GPU tests of the runlmc.linalg mirror classes, restating the reference's
MatrixTestBase pattern (linalg/test_matrix_base.py:33-47): matvec(arange+1) and
matmat against the dense as_numpy(), at the reference's tolerance (1e-6) and at
a much tighter relative one."""
import numpy as np
import pytest
import scipy.linalg as la

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def check_against_dense(mats):
    for M in mats:
        dense = M.as_numpy()
        x = np.arange(dense.shape[1]) + 1.0
        got = M.matvec(x)
        np.testing.assert_allclose(got, dense.dot(x), rtol=1e-6, atol=1e-6, err_msg=str(M))
        assert rel_err(got, dense.dot(x)) < 1e-11 or np.linalg.norm(dense.dot(x)) == 0
        X = np.arange(dense.shape[1] * 2).reshape(-1, 2).astype(float)
        np.testing.assert_allclose(M.matmat(X), dense.dot(X), rtol=1e-6, atol=1e-6, err_msg=str(M))
        # as_linear_operator round trip
        np.testing.assert_allclose(M.as_linear_operator().matvec(x), dense.dot(x), rtol=1e-6, atol=1e-6)


def rpsd(rs, n):
    A = rs.randint(-10, 10, (n, n))
    A = (A + A.T).astype(np.float64)
    A += np.diag(np.fabs(A).sum(axis=1) + 1)
    return A


def toep_eig(e, mult):
    out = np.ones(mult + 1) * 1 - e
    out[0] = 1
    return out


def test_toeplitz_examples():
    from runlmc_b200.linalg import Toeplitz
    rs = np.random.RandomState(3)
    exp = lambda n: np.exp(-rs.rand() * np.arange(n))  # noqa: E731
    tops = [[1], [1, 0], [1, 1], [0, 0], [1, -1], [3.5] + [0.999] * 5 + [0] * 110,
            toep_eig(1e-6 / 2, 5), toep_eig(1e-6, 5), toep_eig(2e-6, 5),
            (np.arange(10) + 1)[::-1], exp(10), exp(50), exp(100), exp(3000)]
    mats = [Toeplitz(np.array(t, dtype=float)) for t in tops]
    check_against_dense(mats)
    for t in mats:
        np.testing.assert_array_equal(t.as_numpy(), la.toeplitz(t.top))


def test_toeplitz_bttb_golden():
    from runlmc_b200.linalg import Toeplitz, BTTB
    g = load_golden('linalg')
    for i in range(int(g['toep_count'])):
        got = Toeplitz(g['toep_top_%d' % i]).matvec(g['toep_x_%d' % i])
        y = g['toep_y_%d' % i]
        np.testing.assert_allclose(got, y, rtol=1e-11, atol=1e-11 * max(1.0, np.abs(y).max()))
    for i in range(int(g['bttb_count'])):
        got = BTTB(g['bttb_top_%d' % i], g['bttb_shape_%d' % i]).matvec(g['bttb_x_%d' % i])
        y = g['bttb_y_%d' % i]
        np.testing.assert_allclose(got, y, rtol=1e-11, atol=1e-11 * max(1.0, np.abs(y).max()))


def test_bttb_examples():
    from runlmc_b200.linalg import BTTB
    rs = np.random.RandomState(4)
    shapes = [(1,), (3,), (2, 3), (10,), (100,), (2, 3, 4), (33, 5), (4, 70), (5, 6, 7), (9000,)]
    mats = []
    for sh in shapes:
        n = int(np.prod(sh))
        if n <= 400:
            mats.append(BTTB(np.arange(n).astype(float), sh))
        mats.append(BTTB(rs.rand(n), sh))
    check_against_dense([m for m in mats if m.shape[0] <= 400])
    # larger ones: against the oracle's FFT path
    from oracle import lmc_oracle as orc
    for m in mats:
        if m.shape[0] > 400:
            x = rs.randn(m.shape[0])
            want = orc.bttb_matvec(orc.bttb_spectrum(m.top, m._sizes), m._sizes, x)
            assert rel_err(m.matvec(x), want) < 1e-12


def test_kronecker_examples():
    from functools import reduce
    from runlmc_b200.linalg import Kronecker, NumpyMatrix, Toeplitz, Matrix
    rs = np.random.RandomState(5)
    up = lambda x: np.diag(np.arange(x) + 1.0)  # noqa: E731
    down = lambda x: up(x)[::-1, ::-1]  # noqa: E731
    exp = lambda n: np.exp(-rs.rand() * np.arange(n))  # noqa: E731
    raw = [
        [up(1), down(1)], [up(3), down(2)], [up(2), down(3)], [la.hilbert(3), la.hilbert(3)],
        [rpsd(rs, 3), np.identity(2)], [rpsd(rs, 2), rpsd(rs, 3)],
        [up(3), Toeplitz(np.arange(10)[::-1] + 1.0)], [Toeplitz(exp(30)), rpsd(rs, 5)],
        [rpsd(rs, 100), rpsd(rs, 5)], [np.identity(2), np.identity(3) * 1e-3],
        [rpsd(rs, 5), Toeplitz(exp(10))], [rpsd(rs, 5), Toeplitz(exp(100))],
        [Toeplitz(exp(10)), Toeplitz(exp(10))], [rs.rand(2, 2) for _ in range(4)],
        [up(2), down(2), up(2)], [rs.rand(2, 3), up(1)], [rs.rand(2, 3), rs.rand(3, 2)],
        [rs.rand(4, 3), rs.rand(5, 2), rs.rand(1, 2)]]
    raw = [[x if isinstance(x, Matrix) else NumpyMatrix(x) for x in ls] for ls in raw]
    mats = [reduce(Kronecker, ls) for ls in raw]
    for k, ls in zip(mats, raw):
        dense = reduce(np.kron, [x.as_numpy() for x in ls])
        assert k.shape == dense.shape
        np.testing.assert_array_equal(k.as_numpy(), dense)
    check_against_dense(mats)
    g = load_golden('linalg')
    got = Kronecker(NumpyMatrix(g['kron_A']), Toeplitz(g['kron_top'])).matvec(g['kron_x'])
    assert rel_err(got, g['kron_y']) < 1e-12


def test_sum_diag_block_examples():
    from runlmc_b200.linalg import (SumMatrix, Kronecker, NumpyMatrix, Toeplitz, Diag, Matrix, BlockDiag,
                                    SymmSquareBlockMatrix, Identity, Composition)
    rs = np.random.RandomState(6)
    exp = lambda n: np.exp(-rs.rand() * np.arange(n))  # noqa: E731
    gen = lambda ms: SumMatrix([m if isinstance(m, Matrix) else NumpyMatrix(m) for m in ms])  # noqa: E731
    examples = [
        [np.diag(np.arange(3) + 1.0), np.diag(np.arange(3) + 1.0)[::-1, ::-1], np.diag(np.ones(3))],
        [rpsd(rs, 3), rpsd(rs, 3), np.diag(rs.rand(3))],
        [la.toeplitz(toep_eig(1e-3 * i, 5)) for i in range(1, 4)] + [np.diag(1e-3 * (1 + rs.rand(6)))],
        [Toeplitz(exp(5)) for _ in range(5)] + [np.diag(np.ones(5) * 1e-4)],
        [Kronecker(NumpyMatrix(rpsd(rs, 2)), Toeplitz(exp(5))) for _ in range(5)] + [np.diag(np.ones(10) * 1e-4)],
        [Kronecker(NumpyMatrix(rpsd(rs, 2)), Kronecker(NumpyMatrix(rpsd(rs, 2)), NumpyMatrix(rpsd(rs, 10))))
         for _ in range(2)] + [np.diag(np.ones(40) * 1e-4)],
        [rpsd(rs, 100) for _ in range(10)] + [np.diag(rs.rand(100))],
        [rs.rand(2, 3), rs.rand(2, 3)], [rs.rand(3, 4) for _ in range(4)]]
    check_against_dense([gen(e) for e in examples])
    g = load_golden('linalg')
    mats = [Kronecker(NumpyMatrix(A), Toeplitz(t)) for A, t in zip(g['sum_As'], g['sum_tops'])]
    mats.append(Diag(g['sum_diag']))
    assert rel_err(SumMatrix(mats).matvec(g['sum_x']), g['sum_y']) < 1e-12
    # Diag / BlockDiag / SymmSquareBlockMatrix / Identity / Composition
    check_against_dense([Diag(rs.rand(7)), Identity(4)])
    bd = BlockDiag([NumpyMatrix(rs.rand(2, 3)), Toeplitz(exp(4)), NumpyMatrix(rs.rand(3, 1))])
    check_against_dense([bd])
    t = [[Toeplitz(exp(6)) for _ in range(3)] for _ in range(3)]
    for i in range(3):
        for j in range(i):
            t[i][j] = t[j][i]
    check_against_dense([SymmSquareBlockMatrix(t)])
    comp = Composition([NumpyMatrix(rs.rand(3, 6)), Toeplitz(exp(6)), NumpyMatrix(rs.rand(6, 2))])
    dense = comp.mats[0].as_numpy().dot(comp.mats[1].as_numpy()).dot(comp.mats[2].as_numpy())
    x = rs.randn(2)
    assert rel_err(comp.matvec(x), dense.dot(x)) < 1e-12
    w = Matrix.wrap((3, 3), lambda v: 2 * v)
    np.testing.assert_allclose(w.matvec(np.arange(3.0)), 2 * np.arange(3.0))
