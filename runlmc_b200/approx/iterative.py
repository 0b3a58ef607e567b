"""Mirror of runlmc/approx/iterative.py: Krylov solves of K x = y."""
import ctypes
import logging

import numpy as np

from .. import _native as nat
from .. import device as dev

_LOG = logging.getLogger(__name__)

_GENERIC_CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p)


def fused_of(K):
    """The fused device operator behind K, if K was recognised as a SKI-LMC tree."""
    return getattr(K, '_fused', None)


def solve_block(K, RHS, tol=1e-4, maxiter=None, check_every=100, minres=True):
    """Batched solve of the rows of RHS with MINRES (or scipy's cg recurrence if not `minres`).
    Returns (X, iters, resid, istop)."""
    RHS = nat.as_f64(RHS)
    if RHS.ndim == 1:
        RHS = RHS.reshape(1, -1)
    fused = fused_of(K)
    if fused is not None:
        solver = fused.minres if minres else fused.cg
        return solver(RHS, tol=tol, maxiter=maxiter, check_every=check_every)
    # arbitrary operator tree: same device MINRES, the product is a callback
    # into the tree's own device kernels
    torch = nat.require_cuda()
    P, n = RHS.shape
    rhs_d = dev.to_device(RHS)
    x_d = dev.empty((P, n))
    s_in = dev.empty((P, n))
    s_out = dev.empty((P, n))

    def apply(_ctx):
        try:
            s_out.copy_(K._apply_dev(s_in))
            return 0
        except Exception:  # pragma: no cover - surfaced as an error code
            _LOG.exception('operator product failed inside MINRES')
            return 3

    cb = _GENERIC_CB(apply)
    iters = np.zeros(P, dtype=np.int32)
    resid = np.zeros(P, dtype=np.float64)
    istop = np.zeros(P, dtype=np.int32)
    entry = nat.lib.lmc_minres_generic if minres else nat.lib.lmc_cg_generic
    nat.check(entry(
        ctypes.cast(cb, ctypes.c_void_p), None, n, dev.ptr(s_in), dev.ptr(s_out), dev.ptr(rhs_d), n, P, dev.ptr(x_d), float(tol),
        int(n if maxiter is None else maxiter), int(check_every), nat.host_ptr(iters),
        nat.host_ptr(resid), nat.host_ptr(istop), dev.stream()))
    return x_d.cpu().numpy(), iters, resid, istop


class Iterative:
    """Target solve() tolerance. Only errors > tol reported."""

    @staticmethod
    def solve(K, y, verbose=False, minres=True, tol=1e-4):
        """Solves K x = y with MINRES (or, with minres=False, scipy's cg) exactly as the
        reference wrapper does (iterative.py:24-62): rtol = min(1e-10, tol), maxiter = n,
        true-residual early termination every 100 iterations; never raises on
        non-convergence, logs instead.

        :return: x, and (iterations, error) too if verbose"""
        if getattr(K, 'preconditioner', None) is not None:
            # the reference forwards it to scipy as M (iterative.py:47) but nothing in runlmc ever sets
            # it; the device solvers implement M = I only, so say so instead of ignoring it
            raise NotImplementedError('preconditioned solves are not part of the accelerated path')
        y = np.asarray(y, dtype=np.float64)
        X, iters, resid, istop = solve_block(K, y.reshape(1, -1), tol=tol, minres=minres)
        Iterative.report(K, resid, istop, tol, minres)
        if verbose:
            return X[0], int(iters[0]), float(resid[0])
        return X[0]

    @staticmethod
    def report(K, resid, istop, tol, minres=True):
        """The reference never raises on non-convergence, it logs (iterative.py:55-58)."""
        n = K.shape[0]
        for r, st in zip(resid, istop):
            exhausted = st == 6 if minres else st not in (0, 10)
            if r > tol or exhausted:
                _LOG.critical('MINRES (n = %d) did not converge in n iterations.'
                              ' Reconstruction error %e', n, r)
