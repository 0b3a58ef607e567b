"""Mirror of runlmc/lmc/likelihood.py (the SKI approximation part).

`ApproxLMCLikelihood` keeps the reference's constructor, attributes and the
four gradient-family methods.  With a fused operator all hyper-parameter
gradients come from ONE device pass (lmc_grad_grams) plus a D x D chain rule on
the host; otherwise each dK operator is applied through the device-backed
linalg tree exactly as the reference loops do (likelihood.py:48-96)."""
import numpy as np

from ..approx.ski import SKI
from ..approx.iterative import fused_of
from ..linalg.diag import Diag
from ..linalg.bttb import BTTB
from ..linalg.kronecker import Kronecker
from ..linalg.numpy_matrix import NumpyMatrix
from ..fused import assemble_gradients


class LMCLikelihood:
    """Gradient families of an LMC log-likelihood in terms of dK/dtheta operators (the role of reference
    likelihood.py:20-96).  A subclass says how a derivative of a coregionalisation matrix, of a kernel or of the
    noise becomes an operator (`_dKdt_from_dAdt`, `_dKdts_from_dKqdts`, `_dKdt_from_dEpsdt`) and how an operator
    becomes a scalar (`_dLdt_from_dKdt`); everything here is the chain rule over the parameter layout
    B_q = A_q^T A_q + diag(kappa_q)."""

    def __init__(self, functional_kernel, Ys):
        self.functional_kernel = functional_kernel
        self.y = np.hstack(Ys)
        self.lens = [len(Y) for Y in Ys]

    def _dKdt_from_dAdt(self, dAdt, q):
        raise NotImplementedError

    def _dKdts_from_dKqdts(self, A, q):
        raise NotImplementedError

    def _dKdt_from_dEpsdt(self, dEpsdt):
        raise NotImplementedError

    def _dLdt_from_dKdt(self, dKdt):
        raise NotImplementedError

    def alpha(self):
        raise NotImplementedError

    def _coreg_partial(self, dB, q):
        return self._dLdt_from_dKdt(self._dKdt_from_dAdt(dB, q))

    def coreg_vec_gradients(self):
        """d/dA_q[r, j]: dB_q = e_j a_r^T + a_r e_j^T."""
        D = self.functional_kernel.D
        out = []
        for q, A in enumerate(self.functional_kernel.coreg_vecs):
            A = np.atleast_2d(A)
            g = np.empty(A.shape)
            for r, j in np.ndindex(*A.shape):
                dB = np.zeros((D, D))
                dB[j, :] += A[r]
                dB[:, j] += A[r]
                g[r, j] = self._coreg_partial(dB, q)
            out.append(g)
        return out

    def coreg_diags_gradients(self):
        """d/dkappa_q[i]: dB_q = e_i e_i^T."""
        D, Q = self.functional_kernel.D, self.functional_kernel.Q
        units = np.eye(D)
        return [np.array([self._coreg_partial(np.outer(units[i], units[i]), q) for i in range(D)])
                for q in range(Q)]

    def kernel_gradients(self):
        return [[self._dLdt_from_dKdt(dK) for dK in self._dKdts_from_dKqdts(B, q)]
                for q, B in enumerate(self.functional_kernel.coreg_mats())]

    def noise_gradient(self):
        units = np.eye(self.functional_kernel.D)
        return np.array([self._dLdt_from_dKdt(self._dKdt_from_dEpsdt(e)) for e in units])


class ApproxLMCLikelihood(LMCLikelihood):
    def __init__(self, functional_kernel, grid_kern, grid_dists, interpolants, Ys, deriv):
        super().__init__(functional_kernel, Ys)
        self.grid_dists = grid_dists
        self._materialized = None
        self._materialized_grads = None
        self.K = grid_kern
        self.deriv = deriv.generate(self.K, self.y)
        self.interpolants = interpolants
        self._fused_grads = None
        # like the reference, everything the gradients need is fixed at construction: the handle behind K is
        # shared by later operators with other hyper-parameters, so the Gram stage runs now, on K's own
        self._fused()

    # The reference evaluates both on construction (likelihood.py:103-108); here they are evaluated
    # on first use, because with device-evaluated kernels (lmc_op_set_kernels) the fused gradient
    # path never needs them on the host.
    @property
    def materialized_kernels(self):
        if self._materialized is None:
            self._materialized = [BTTB(d.ravel(), d.shape)
                                  for d in self.functional_kernel.eval_kernels(self.grid_dists)]
        return self._materialized

    @property
    def materialized_grads(self):
        if self._materialized_grads is None:
            self._materialized_grads = self.functional_kernel.eval_kernel_gradients(self.grid_dists)
        return self._materialized_grads

    # -- fused path -------------------------------------------------------
    def _fused(self):
        fused = fused_of(self.K)
        if fused is None or len(self.functional_kernel.active_dims) != 1:
            return None
        if self._fused_grads is None:
            fk = self.functional_kernel
            d = self.deriv
            if fused.kernels_on_device:
                extra = None          # derivative tops are evaluated on the device too
                counts = [len(k.param_values()) for k in fk._kernels]
            else:
                extra = [t for ts in self.materialized_grads for t in ts]
                counts = [len(t) for t in self.materialized_grads]
            a, R, Rinv = d._dev()          # already on the device when the service solved with a fused operator
            quad, trace, nquad, ntrace = fused.grad_grams_device(a[0], R if len(R) else None,
                                                                  Rinv if len(R) else None, extra)
            self._fused_grads = assemble_gradients(
                fk.coreg_vecs, fk.coreg_mats(), counts, d._n_it, quad, trace, nquad, ntrace)
        return self._fused_grads

    def coreg_vec_gradients(self):
        f = self._fused()
        return f[0] if f is not None else super().coreg_vec_gradients()

    def coreg_diags_gradients(self):
        f = self._fused()
        return f[1] if f is not None else super().coreg_diags_gradients()

    def kernel_gradients(self):
        f = self._fused()
        return f[2] if f is not None else super().kernel_gradients()

    def noise_gradient(self):
        f = self._fused()
        return f[3] if f is not None else super().noise_gradient()

    # -- generic path (likelihood.py:112-131) --------------------------------
    def _ski(self, q, X):
        ad = self.functional_kernel.get_active_dims(q)
        return SKI(X, *self.interpolants[ad])

    def _dKdt_from_dAdt(self, dAdt, q):
        return self._ski(q, Kronecker(NumpyMatrix(dAdt), self.materialized_kernels[q]))

    def _dKdts_from_dKqdts(self, A, q):
        for dKqdt in self.materialized_grads[q]:
            yield self._ski(q, Kronecker(NumpyMatrix(A), BTTB(dKqdt.ravel(), dKqdt.shape)))

    def _dKdt_from_dEpsdt(self, dEpsdt):
        return Diag(np.repeat(dEpsdt, self.lens))

    def _dLdt_from_dKdt(self, dKdt):
        return self.deriv.derivative(dKdt)

    def alpha(self):
        return self.deriv.alpha
