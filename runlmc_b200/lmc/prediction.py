"""Predictive mean and variance of the SKI-LMC model on the fused device operator.

Mirrors InterpolatedLLGP._raw_predict and its three variance modes
(reference runlmc/models/interpolated_llgp.py:293-397).  The reference fans the
variance solves out one right-hand side per task through
``pool.starmap(Iterative.solve, ...)`` (:373-380, :390-397); here they are the
same multi-RHS MINRES the gradient path uses, in blocks of `block` columns, and
every product runs in liblmc_b200.so.  One active-dimension group (the case the
fused operator covers).
"""
import numpy as np
import scipy.spatial.distance as _dist

from .. import _native as nat
from ..fused import FusedLMC


def kernel_from_indices(Xs, Zs, functional_kernel):
    """Dense exact LMC cross-covariance between the points of `Xs` and `Zs`
    (lists of per-output inputs): ExactLMCLikelihood.kernel_from_indices,
    lmc/likelihood.py:183-203.  Host side: it only builds right-hand sides."""
    fk = functional_kernel
    rlens, clens = [len(X) for X in Xs], [len(Z) for Z in Zs]
    X = np.vstack([np.asarray(x, dtype=np.float64).reshape(len(x), -1) for x in Xs])
    Z = np.vstack([np.asarray(z, dtype=np.float64).reshape(len(z), -1) for z in Zs])
    dists = {ad: _dist.cdist(X[:, list(ad)], Z[:, list(ad)]) for ad in fk.active_dims}
    Kqs = fk.eval_kernels(dists)
    rb = np.concatenate([[0], np.cumsum(rlens)])
    cb = np.concatenate([[0], np.cumsum(clens)])
    K = np.zeros((len(X), len(Z)))
    for A, Kq in zip(fk.coreg_mats(), Kqs):
        Kq = np.array(Kq, dtype=np.float64, copy=True)
        for i in range(len(rlens)):
            for j in range(len(clens)):
                Kq[rb[i]:rb[i + 1], cb[j]:cb[j + 1]] *= A[i, j]
        K += Kq
    return K


class Predictor:
    """:param fused: FusedLMC over the training inputs, parameters set
    :param functional_kernel: the kernel container (coreg_vecs, coreg_diags, noise,
        eval_kernels, coreg_mats, active_dims)
    :param Xs_train: training inputs per output (only the on-the-fly mode needs them)
    :param grids: the per-axis grids the operator was built on
    :param alpha: K^-1 y in the caller's point order (numpy)
    :param tol: MINRES tolerance of the variance solves (reference default 1e-4)
    :param block: right-hand sides per batched solve
    """

    def __init__(self, fused, functional_kernel, Xs_train, grids, alpha, tol=1e-4, block=128):
        self.fused, self.fk = fused, functional_kernel
        self.Xs_train = Xs_train
        self.grids = [np.asarray(g, dtype=np.float64) for g in grids]
        self.alpha = np.asarray(alpha, dtype=np.float64)
        self.tol, self.block = tol, int(block)
        self._grid_alpha = None
        self._nu = None

    # ---- pieces -----------------------------------------------------------
    def native_variance(self):
        """A-priori variance of one point of every output (interpolated_llgp.py:304-316)."""
        fk = self.fk
        coregs = np.column_stack([np.square(np.atleast_2d(a)).sum(axis=0) for a in fk.coreg_vecs])
        coregs = coregs + np.column_stack(fk.coreg_diags)
        zero = {ad: np.zeros(1) for ad in fk.active_dims}
        k0 = np.array([np.asarray(k).reshape(-1)[0] for k in fk.eval_kernels(zero)])
        return coregs.dot(k0).reshape(-1) + np.asarray(fk.noise)

    def _test_op(self, Xs):
        return FusedLMC(Xs, self.grids)

    def grid_alpha(self):
        """K_UU W^T alpha on the device (interpolated_llgp.py:293-300)."""
        if self._grid_alpha is None:
            torch = nat.require_cuda()
            a = torch.as_tensor(self.alpha.reshape(1, -1), device='cuda')
            self._grid_alpha = self.fused.grid_mvm_device(self.fused.to_grid_device(a))
        return self._grid_alpha

    def mean(self, Xs):
        """W* (K_UU W^T alpha) (interpolated_llgp.py:334-338), flat over outputs."""
        return self._test_op(Xs).from_grid_device(self.grid_alpha())[0].cpu().numpy()

    def nu(self):
        """nu_i = e_i' K_UX K^-1 K_XU e_i for all D*m grid entries (_precomputed_nu,
        interpolated_llgp.py:358-382), `block` unit vectors per batched solve."""
        if self._nu is not None:
            return self._nu
        torch = nat.require_cuda()
        f = self.fused
        Dm = f.D * f.m
        nu = np.empty(Dm)
        for i0 in range(0, Dm, self.block):
            i1 = min(Dm, i0 + self.block)
            E = torch.zeros((i1 - i0, Dm), dtype=torch.float64, device='cuda')
            idx = torch.arange(i0, i1, device='cuda')
            E[torch.arange(i1 - i0, device='cuda'), idx] = 1.0
            rhs = f.from_grid_device(f.grid_mvm_device(E))                  # K_XU e_i = W K_UU e_i
            X, _, _, _ = f.minres_device(rhs, tol=self.tol)
            back = f.grid_mvm_device(f.to_grid_device(X))                   # K_UX x = K_UU W^T x
            nu[i0:i1] = back[torch.arange(i1 - i0, device='cuda'), idx].cpu().numpy()
        self._nu = nu
        return nu

    def var_precompute(self, Xs):
        """Explained variance W* nu (_var_predict_precompute, interpolated_llgp.py:384-388)."""
        torch = nat.require_cuda()
        g = torch.as_tensor(self.nu().reshape(1, -1), device='cuda')
        return self._test_op(Xs).from_grid_device(g)[0].cpu().numpy()

    def var_on_the_fly(self, Xs):
        """Explained variance diag(K_*X K^-1 K_X*) with one solve per test point
        (_var_predict_on_the_fly, interpolated_llgp.py:390-397), `block` points per batched solve."""
        torch = nat.require_cuda()
        Kx = kernel_from_indices(Xs, self.Xs_train, self.fk)
        out = np.empty(len(Kx))
        for i0 in range(0, len(Kx), self.block):
            rhs = torch.as_tensor(np.ascontiguousarray(Kx[i0:i0 + self.block]), device='cuda')
            X, _, _, _ = self.fused.minres_device(rhs, tol=self.tol)
            out[i0:i0 + self.block] = (rhs * X).sum(dim=1).cpu().numpy()
        return out

    # ---- the reference's entry point ---------------------------------------
    def predict(self, Xs, mode='on-the-fly'):
        """_raw_predict (interpolated_llgp.py:324-348): per-output lists (means, variances);
        variances are native minus explained, clipped at zero."""
        if mode not in ('on-the-fly', 'precompute'):
            raise ValueError('Variance prediction mode {} should be one of {}'.format(
                mode, ['on-the-fly', 'precompute']))
        Xs = [np.asarray(X, dtype=np.float64).reshape(len(X), -1) for X in Xs]
        lens = [len(X) for X in Xs]
        mean = self.mean(Xs)
        native = np.repeat(self.native_variance(), lens)
        explained = self.var_precompute(Xs) if mode == 'precompute' else self.var_on_the_fly(Xs)
        var = native - explained
        var[var < 0] = 0
        ends = np.add.accumulate(lens)[:-1]
        return np.split(mean, ends), np.split(var, ends)
