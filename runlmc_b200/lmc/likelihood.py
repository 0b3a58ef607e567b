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
    def __init__(self, functional_kernel, Ys):
        self.functional_kernel = functional_kernel
        self.y = np.hstack(Ys)
        self.lens = list(map(len, Ys))

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

    def coreg_vec_gradients(self):
        fk = self.functional_kernel
        grads = []
        for q, a in enumerate(fk.coreg_vecs):
            g = np.zeros(np.shape(a))
            for i, ai in enumerate(a):
                for j in range(fk.D):
                    dA = np.zeros((fk.D, fk.D))
                    dA[j] += ai
                    dA.T[j] += ai
                    g[i, j] = self._dLdt_from_dKdt(self._dKdt_from_dAdt(dA, q))
            grads.append(g)
        return grads

    def coreg_diags_gradients(self):
        fk = self.functional_kernel
        grads = []
        for q in range(fk.Q):
            g = np.zeros(fk.D)
            for i in range(fk.D):
                E = np.zeros((fk.D, fk.D))
                E[i, i] = 1
                g[i] = self._dLdt_from_dKdt(self._dKdt_from_dAdt(E, q))
            grads.append(g)
        return grads

    def kernel_gradients(self):
        grads = []
        for q, A in enumerate(self.functional_kernel.coreg_mats()):
            grads.append([self._dLdt_from_dKdt(dK) for dK in self._dKdts_from_dKqdts(A, q)])
        return grads

    def noise_gradient(self):
        D = self.functional_kernel.D
        g = np.zeros(D)
        for i in range(D):
            e = np.zeros(D)
            e[i] = 1
            g[i] = self._dLdt_from_dKdt(self._dKdt_from_dEpsdt(e))
        return g


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
            quad, trace, nquad, ntrace = fused.grad_grams(
                d.alpha, np.asarray(d._rs, dtype=np.float64), np.asarray(d._inv_rs), extra)
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
