"""Mirror of runlmc/lmc/stochastic_deriv.py: Hutchinson-probe derivative service."""
import numpy as np

from ..approx.iterative import Iterative, JacobiPreconditioner, fused_of, solve_block
from .. import _native as nat
from .. import device as dev


class Derivative:
    """A log-likelihood derivative in the two pieces every estimator provides: the data-fit term
    alpha' dK alpha and the log-determinant term tr(K^-1 dK); dL/dtheta is half their difference
    (reference lmc/derivative.py:5-12)."""

    def d_normal_quadratic(self, dKdt):
        raise NotImplementedError

    def d_logdet_K(self, dKdt):
        raise NotImplementedError

    def derivative(self, dKdt):
        fit, logdet = self.d_normal_quadratic(dKdt), self.d_logdet_K(dKdt)
        return (fit - logdet) / 2


class Metrics:
    """What a run records per optimiser step when asked to (reference lmc/metrics.py:4-10): solver
    iterations and residuals from the derivative service, gradient norms / errors and the
    log-likelihood from the model."""
    FIELDS = ('iterations', 'grad_norms', 'grad_error', 'solv_error', 'log_likely')

    def __init__(self):
        for name in self.FIELDS:
            setattr(self, name, [])


class StochasticDerivService:
    """:param metrics: a Metrics instance or None, :param pool: pool for operator
    trees the fused path does not cover, :param n_it: number of probes,
    :param tol: solver tolerance (stochastic_deriv.py:27-31).
    :param logdet_steps: (extension, default off) keep this many Lanczos steps of every probe solve and
        attach the stochastic-Lanczos-quadrature estimate of log det K to the result (``log_det_K``,
        ``log_det_K_stderr``; approx/logdet.py) -- the reference only has a dense Cholesky for it
        (models/interpolated_llgp.py:262-276).  Fused operators only.
    :param device_probes: (extension, default off) draw the Rademacher probes on the device (torch's
        generator, seeded with ``torch.manual_seed``) instead of from the global numpy RNG: no 8 n N bytes of
        host random numbers, no upload.  Same estimator, a different random stream than the reference's."""

    def __init__(self, metrics, pool, n_it, tol, logdet_steps=0, device_probes=False):
        self.metrics = metrics
        self._pool = pool
        self._n_it = n_it
        self._tol = tol
        self._logdet_steps = logdet_steps
        self._device_probes = device_probes

    def generate(self, K, y, rs=None):
        """Draw n_it Rademacher probes from the GLOBAL numpy RNG exactly like the
        reference (stochastic_deriv.py:35) -- or take host-supplied `rs` -- and
        solve the n_it + 1 systems as one multi-RHS MINRES on the device.  With a fused operator the
        right-hand sides are uploaded once and the solutions stay on the device for the gradient
        contractions (host copies are made only if someone reads them)."""
        n = K.shape[0]
        fused = fused_of(K)
        M = getattr(K, 'preconditioner', None)      # honoured like Iterative.solve does (iterative.py:47)
        jacobi = isinstance(M, JacobiPreconditioner) and M.serves(K)
        if fused is not None and (M is None or jacobi):
            return self._generate_fused(fused, y, rs, n, 'jacobi' if jacobi else None)
        if rs is None:
            rs = np.random.randint(0, 2, (self._n_it, n)) * 2 - 1
        rs = np.asarray(rs)
        RHS = np.vstack([np.asarray(y, dtype=np.float64).reshape(1, -1), rs.astype(np.float64)])
        X, iters, resid, _ = solve_block(K, RHS, tol=self._tol, preconditioner=M)
        self._record(iters, resid)
        return StochasticDeriv(X[0], rs, list(X[1:]), self._n_it)

    def _record(self, iters, resid):
        if self.metrics is not None:
            self.metrics.iterations.append(np.mean(iters))
            self.metrics.solv_error.append(np.mean(resid))

    def _generate_fused(self, fused, y, rs, n, precond):
        torch = nat.require_cuda()
        N = self._n_it if rs is None else len(rs)
        RHS = torch.empty((N + 1, n), dtype=torch.float64, device='cuda')
        RHS[0] = torch.as_tensor(np.asarray(y, dtype=np.float64).reshape(-1))
        if rs is None and self._device_probes:
            RHS[1:] = torch.randint(0, 2, (N, n), device='cuda', dtype=torch.int8) * 2 - 1
        else:
            if rs is None:
                rs = np.random.randint(0, 2, (self._n_it, n)) * 2 - 1
            rs = np.asarray(rs)
            RHS[1:] = torch.as_tensor(rs).to('cuda')      # the cast to float64 happens on the device
        logdet = None
        if self._logdet_steps and precond is None:
            from ..approx.logdet import quadrature_terms
            X, iters, resid, _, tri, beta1 = fused.minres_lanczos_device(RHS, self._logdet_steps, tol=self._tol)
            logdet = quadrature_terms(tri[1:], beta1[1:], iters[1:])      # the probes, not y
        else:
            X, iters, resid, istop = fused.minres_device(RHS, tol=self._tol, precond=precond)
            if np.any(istop == 9):
                raise ValueError('indefinite preconditioner')
        self._record(iters, resid)
        deriv = StochasticDeriv._on_device(X[:1], RHS[1:], X[1:], N, rs)
        if logdet is not None and len(logdet):
            deriv.log_det_K = float(np.mean(logdet))
            deriv.log_det_K_stderr = float(np.std(logdet, ddof=1) / np.sqrt(len(logdet))) if len(logdet) > 1 \
                else float('nan')
        return deriv

    def _concurrent_solve(self, ls):
        return self._pool.starmap(Iterative.solve, ls)


class StochasticDeriv(Derivative):
    """Derivatives from alpha = K^-1 y and the probe solves K^-1 r_i
    (stochastic_deriv.py:55-78)."""

    def __init__(self, alpha, rs, inv_rs, n_it):
        self.alpha = alpha
        self._rs_host = rs
        self._inv_rs_host = inv_rs
        self._n_it = n_it
        self._dev_cache = None
        self.log_det_K = None            # set by a service created with logdet_steps > 0
        self.log_det_K_stderr = None

    @classmethod
    def _on_device(cls, alpha_d, R_d, Rinv_d, n_it, rs_host=None):
        """alpha [1, n], probes and probe solves [N, n] as CUDA tensors (views of the solver's blocks)."""
        self = cls(alpha_d[0].cpu().numpy(), rs_host, None, n_it)
        self._dev_cache = (alpha_d.contiguous(), R_d.contiguous(), Rinv_d.contiguous())
        return self

    # the reference's attributes, copied to the host only when read
    @property
    def _rs(self):
        if self._rs_host is None and self._dev_cache is not None:
            self._rs_host = self._dev_cache[1].cpu().numpy()
        return self._rs_host

    @property
    def _inv_rs(self):
        if self._inv_rs_host is None and self._dev_cache is not None:
            self._inv_rs_host = list(self._dev_cache[2].cpu().numpy())
        return self._inv_rs_host

    def _dev(self):
        if self._dev_cache is None:
            self._dev_cache = (dev.to_device(np.asarray(self.alpha).reshape(1, -1)),
                               dev.to_device(np.asarray(self._rs, dtype=np.float64)),
                               dev.to_device(np.asarray(self._inv_rs, dtype=np.float64)))
        return self._dev_cache

    @staticmethod
    def _dot(A, B):
        out = np.zeros(1)
        nat.check(nat.lib.lmc_block_dot(dev.ptr(A), A.shape[1], dev.ptr(B), B.shape[1], A.shape[1],
                                        A.shape[0], nat.host_ptr(out), dev.stream()))
        return float(out[0])

    def d_normal_quadratic(self, dKdt):
        a, _, _ = self._dev()
        return self._dot(a, dKdt._apply_dev(a))

    def d_logdet_K(self, dKdt):
        _, R, Rinv = self._dev()
        return self._dot(Rinv, dKdt._apply_dev(R)) / self._n_it
