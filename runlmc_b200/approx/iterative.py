"""Mirror of runlmc/approx/iterative.py: Krylov solves of K x = y."""
import ctypes
import logging

import numpy as np

from .. import _native as nat
from .. import device as dev
from ..linalg.diag import Diag

_LOG = logging.getLogger(__name__)

_GENERIC_CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p)


def fused_of(K):
    """The fused device operator behind K, if K was recognised as a SKI-LMC tree."""
    return getattr(K, '_fused', None)


class JacobiPreconditioner(Diag):
    """M = diag(K~)^-1 for a fused SKI-LMC operator K (the exact diagonal, computed on the device
    without forming K~: lmc_op_diagonal).  Set ``K.preconditioner = JacobiPreconditioner(K)`` and
    Iterative.solve forwards it as scipy's ``M`` like the reference does (iterative.py:47-50); on the
    fused path the block solver applies it in its own sorted layout (lmc_minres_pre)."""

    def __init__(self, K):
        fused = fused_of(K)
        if fused is None:
            raise ValueError('JacobiPreconditioner needs a fused SKI-LMC operator (gen_grid_kernel)')
        super().__init__(1.0 / fused.diagonal())
        self._K = K

    def serves(self, K):
        return K is self._K and fused_of(K) is not None


def solve_block(K, RHS, tol=1e-4, maxiter=None, check_every=100, minres=True, preconditioner=None):
    """Batched solve of the rows of RHS with MINRES (or scipy's cg recurrence if not `minres`).
    `preconditioner`: scipy's M, any SPD runlmc_b200.linalg Matrix (MINRES only).
    Returns (X, iters, resid, istop)."""
    RHS = nat.as_f64(RHS)
    if RHS.ndim == 1:
        RHS = RHS.reshape(1, -1)
    M = preconditioner
    if M is not None and not minres:
        raise NotImplementedError('the device cg solver has no preconditioned variant; use minres=True')
    if M is not None and tuple(M.shape) != tuple(K.shape):
        raise ValueError('preconditioner shape {} != operator shape {}'.format(M.shape, K.shape))
    fused = fused_of(K)
    if fused is not None and (M is None or (isinstance(M, JacobiPreconditioner) and M.serves(K))):
        if M is not None:
            return fused.minres(RHS, tol=tol, maxiter=maxiter, check_every=check_every, precond='jacobi')
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
    if M is not None:
        def precondition(_ctx):
            try:
                s_out.copy_(M._apply_dev(s_in))
                return 0
            except Exception:  # pragma: no cover - surfaced as an error code
                _LOG.exception('preconditioner product failed inside MINRES')
                return 3

        pcb = _GENERIC_CB(precondition)
        nat.check(nat.lib.lmc_minres_generic_pre(
            ctypes.cast(cb, ctypes.c_void_p), ctypes.cast(pcb, ctypes.c_void_p), None, n, dev.ptr(s_in),
            dev.ptr(s_out), dev.ptr(rhs_d), n, P, dev.ptr(x_d), float(tol), int(n if maxiter is None else maxiter),
            int(check_every), nat.host_ptr(iters), nat.host_ptr(resid), nat.host_ptr(istop), dev.stream()))
        return x_d.cpu().numpy(), iters, resid, istop
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
        # the reference forwards an optional K.preconditioner to scipy as M (iterative.py:47-50)
        M = getattr(K, 'preconditioner', None)
        y = np.asarray(y, dtype=np.float64)
        X, iters, resid, istop = solve_block(K, y.reshape(1, -1), tol=tol, minres=minres, preconditioner=M)
        if minres and istop[0] == 9:
            raise ValueError('indefinite preconditioner')        # what scipy raises (minres.py:259)
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
