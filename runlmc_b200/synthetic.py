"""Synthetic SKI-LMC problems.

Mirrors the recipe of the reference's benchmark driver
(benchmarks/benchlib/bench.py:105-143): uniform inputs/outputs,
truncated-normal coregionalisation vectors, inverse-gamma kappa and noise,
RBF kernels with log-spaced inverse lengthscales.  Grids are built directly
(``linspace(0, 1, m_p)``) so the embedding is the power of two the config
names, rather than through ``autogrid`` which adds 4 points
(approx/interpolation.py:212).

Shared by bench.py and the tests; pure numpy, no device code.
"""
import numpy as np
import scipy.stats

# name -> (D, per-output n, grid sizes, Q, probes)
CONFIGS = {
    'A': dict(D=2, lens=[65, 100], grid=[86], Q=2, N=15),            # README snippet scale
    'B': dict(D=2, lens=[5000] * 2, grid=[1024], Q=2, N=16),         # benchmarks/synth-like
    'C': dict(D=13, lens=[235] * 13, grid=[238], Q=3, N=16),         # fx2007-like
    'D': dict(D=4, lens=[125000] * 4, grid=[8192], Q=3, N=64),       # weather-like, n=500k
    'E': dict(D=10, lens=[100000] * 10, grid=[256, 256], Q=3, N=128),  # 2-D BTTB, n=1M
    # small stand-ins with the same structure, for fast parity tests
    'd_small': dict(D=4, lens=[700, 650, 720, 690], grid=[256], Q=3, N=6),
    'e_small': dict(D=3, lens=[900, 800, 850], grid=[32, 16], Q=3, N=6),
}


class Problem:
    """Plain container: everything both backends need, as numpy arrays."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def n(self):
        return int(sum(self.lens))

    def coreg_mats(self):
        return [a.T.dot(a) + np.diag(k)
                for a, k in zip(self.coreg_vecs, self.coreg_diags)]


def rbf_top(dists, gamma):
    """k(r) = exp(-gamma r^2 / 2) (reference kern/rbf.py:39-40)."""
    return np.exp(-0.5 * np.square(dists) * gamma)


def rbf_top_grad(dists, gamma):
    """dk/dgamma (reference kern/rbf.py:50-54)."""
    sq = np.square(dists)
    return np.exp(-0.5 * sq * gamma) * -0.5 * sq


def make_problem(name=None, seed=1234, eps=0.1, cells_per_lengthscale=None,
                 edge=False, noise_scale=1.0, **override):
    """Build a synthetic problem.

    ``cells_per_lengthscale``: if given, the Q RBF inverse lengthscales are
    chosen so the lengthscales span [c, 4c] grid cells (log-spaced) instead of
    bench.py's logspace(0, 1, Q) -- needed on fine grids so the kernel is
    resolved by the grid and K~ is not numerically rank-deficient.
    ``edge``: draw inputs from U(0, 1) (touching the grid boundary, exercising
    the clamped stencils) instead of U(0.02, 0.98).
    ``noise_scale``: multiplies the drawn noise variances (a low signal-to-noise model: at n = 1M the
    reference's stopping rule, relative to ||A|| ||x||, only reaches an absolute residual of 1e-4 when
    K~ is well conditioned, bench.py's converging gradient row)."""
    cfg = dict(CONFIGS[name]) if name else {}
    cfg.update(override)
    D, lens, grid_sizes, Q, N = (cfg['D'], list(cfg['lens']),
                                 list(cfg['grid']), cfg['Q'], cfg['N'])
    rng = np.random.default_rng(seed)
    ndim = len(grid_sizes)
    lo, hi = (0.0, 1.0) if edge else (0.02, 0.98)
    Xs = [rng.uniform(lo, hi, size=(nd, ndim)) for nd in lens]
    Ys = [rng.uniform(0, 1, size=nd) for nd in lens]
    grids = [np.linspace(0, 1, m) for m in grid_sizes]
    mesh = np.stack(np.meshgrid(*grids, indexing='ij'), axis=-1)
    dists = np.linalg.norm(mesh - mesh.reshape(-1, ndim)[0], axis=-1)

    tn = scipy.stats.truncnorm(-1, 1)
    coreg_vecs = [tn.rvs(size=(1, D), random_state=rng) for _ in range(Q)]
    coreg_diags = [np.reciprocal(rng.gamma(1.0, 1.0, size=D))
                   for _ in range(Q)]
    noise = np.reciprocal(rng.gamma(1 + 1 / eps, 1.0, size=D)) * noise_scale
    if cells_per_lengthscale is None:
        gammas = np.logspace(0, 1, Q)
    else:
        delta = 1.0 / (max(grid_sizes) - 1)
        ells = np.geomspace(4 * cells_per_lengthscale,
                            cells_per_lengthscale, Q) * delta
        gammas = 1.0 / np.square(ells)
    tops = [rbf_top(dists, g) for g in gammas]
    top_grads = [[rbf_top_grad(dists, g)] for g in gammas]
    n = sum(lens)
    probes = (rng.integers(0, 2, size=(N, n)) * 2 - 1).astype(np.float64)
    return Problem(name=name, D=D, Q=Q, N=N, lens=lens, ndim=ndim,
                   grid_sizes=grid_sizes, grids=grids, dists=dists,
                   Xs=Xs, Ys=Ys, y=np.hstack(Ys), coreg_vecs=coreg_vecs,
                   coreg_diags=coreg_diags, noise=noise, gammas=gammas,
                   tops=tops, top_grads=top_grads, probes=probes)
