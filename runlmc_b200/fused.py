"""Python handle of the fused device operator  K~ = W (sum_q B_q (x) T_q) W^T + D.

One `FusedLMC` replaces the whole operator tree the reference assembles in
gen_grid_kernel (runlmc/lmc/grid_kernel.py:49-74).  All arithmetic happens in
liblmc_b200.so; torch is used for device buffers and streams only.
"""
import ctypes

import numpy as np

from . import _native as nat


def _dev(t):
    return ctypes.c_void_p(t.data_ptr())


def _no_handle():
    return None


KERNEL_KINDS = {'RBF': 0, 'Matern32': 1, 'StdPeriodic': 2}   # LMC_KERN_* of include/lmc_b200.h


def kernel_descriptor(k):
    """(kind, inv_lengthscale, period) of a runlmc_b200.kern kernel the device can evaluate, or
    None (then its values are computed on the host and uploaded)."""
    from . import kern as _kern
    for cls in (_kern.RBF, _kern.Matern32, _kern.StdPeriodic):
        if type(k) is cls:
            return KERNEL_KINDS[cls.__name__], float(k.inv_lengthscale), float(getattr(k, 'period', 1.0))
    return None


class FusedLMC:
    """:param Xs: list of D arrays of input points, each [n_d] or [n_d, ndim]
    :param grids: list of ndim equispaced 1-D grids (ndim in {1, 2})
    :param build: 'device' (default): coordinates are uploaded once and the point sort runs on the
        GPU (lmc_op_create_dev); 'host': the host counting sort (lmc_op_create).  Same operator,
        bit for bit.
    """

    def __init__(self, Xs, grids, build='device'):
        self._h = ctypes.c_void_p()
        nat.require_cuda()
        grids = [np.asarray(g, dtype=np.float64) for g in grids]
        ndim = len(grids)
        Xs = [np.asarray(X, dtype=np.float64) for X in Xs]
        if any(X.size != len(X) * ndim for X in Xs):
            raise ValueError('input dimension does not match number of grids')
        Xs = [X.reshape(len(X), ndim) for X in Xs]
        for g in grids:
            if g.ndim != 1:
                raise ValueError('grid dim {} should be 1'.format(g.ndim))
            if g.size < 4:
                raise ValueError('grid size {} must be >=4'.format(g.size))
        self.D = len(Xs)
        self.ndim = ndim
        self.lens = [len(X) for X in Xs]
        self.n = int(sum(self.lens))
        self.grid_sizes = [int(g.size) for g in grids]
        self.m = int(np.prod(self.grid_sizes))
        sizes = nat.as_i32(self.grid_sizes)
        origin = nat.as_f64([g[0] for g in grids])
        delta = nat.as_f64([g[1] - g[0] for g in grids])   # interpolation.py:98
        lens = nat.as_i32(self.lens)
        X = nat.as_f64(np.vstack(Xs)) if self.n else np.zeros((0, ndim))
        if build == 'device' and self.n:
            torch = nat.require_cuda()
            Xd = torch.as_tensor(X, device='cuda')
            nat.check(nat.lib.lmc_op_create_dev(
                ctypes.byref(self._h), self.D, ndim, nat.host_ptr(sizes),
                nat.host_ptr(origin), nat.host_ptr(delta), nat.host_ptr(lens),
                _dev(Xd), nat.current_stream_ptr()))
        elif build in ('host', 'device'):
            nat.check(nat.lib.lmc_op_create(
                ctypes.byref(self._h), self.D, ndim, nat.host_ptr(sizes),
                nat.host_ptr(origin), nat.host_ptr(delta), nat.host_ptr(lens),
                nat.host_ptr(X)))
        else:
            raise ValueError('build should be "device" or "host"')
        self.grid_deltas = [float(d) for d in delta]
        self.kernels_on_device = False
        self.Q = 0
        self.shape = (self.n, self.n)
        self.dtype = np.float64

    def __reduce__(self):
        # device handles do not travel: whatever holds one (e.g. the cache gen_grid_kernel keeps on an
        # interpolant) unpickles to None and is rebuilt on demand
        return (_no_handle, ())

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            try:
                nat.lib.lmc_op_destroy(h)
            except Exception:  # interpreter shutdown
                pass

    # ---- parameters ------------------------------------------------------
    def set_params(self, tops, Bs, noise, coreg_vecs=None, coreg_diags=None):
        """tops: Q arrays of kernel values on the grid (any shape with m
        entries); Bs: Q (D, D) coregionalisation matrices; noise: (D,).
        Optionally the LMC factors with Bs[q] = coreg_vecs[q].T @ coreg_vecs[q]
        + diag(coreg_diags[q]) (functional_kernel.py:280-287), which make the
        spectral mix cheaper."""
        tops = nat.as_f64(np.array([np.asarray(t, dtype=np.float64).ravel()
                                    for t in tops]))
        Bs = nat.as_f64(np.array(Bs))
        noise = nat.as_f64(noise)
        Q = tops.shape[0]
        if tops.shape != (Q, self.m):
            raise ValueError('tops shape {} != {}'.format(tops.shape, (Q, self.m)))
        if Bs.shape != (Q, self.D, self.D):
            raise ValueError('B shape {} != {}'.format(Bs.shape, (Q, self.D, self.D)))
        if noise.shape != (self.D,):
            raise ValueError('noise shape {} != {}'.format(noise.shape, (self.D,)))
        nat.check(nat.lib.lmc_op_set_params(
            self._h, Q, nat.host_ptr(tops), nat.host_ptr(Bs), nat.host_ptr(noise)))
        self.Q = Q
        self.kernels_on_device = False
        self._set_factors(Q, coreg_vecs, coreg_diags)

    def set_kernels(self, kerns, Bs, noise, coreg_vecs=None, coreg_diags=None):
        """Same update with the kernel values evaluated on the device (lmc_op_set_kernels): kerns
        are Q runlmc_b200.kern RBF / Matern32 / StdPeriodic objects; their values on the operator's
        own grid distances never exist on the host."""
        desc = [kernel_descriptor(k) for k in kerns]
        if any(d is None for d in desc):
            raise ValueError('only RBF, Matern32 and StdPeriodic kernels are evaluated on the device')
        self.set_kernel_descriptors(desc, Bs, noise, coreg_vecs, coreg_diags)

    def set_kernel_descriptors(self, desc, Bs, noise, coreg_vecs=None, coreg_diags=None):
        """set_kernels from (kind, inv_lengthscale, period) triples (kernel_descriptor)."""
        kinds = nat.as_i32([d[0] for d in desc])
        params = nat.as_f64([[d[1], d[2]] for d in desc])
        Bs = nat.as_f64(np.array(Bs))
        noise = nat.as_f64(noise)
        Q = len(desc)
        if Bs.shape != (Q, self.D, self.D):
            raise ValueError('B shape {} != {}'.format(Bs.shape, (Q, self.D, self.D)))
        if noise.shape != (self.D,):
            raise ValueError('noise shape {} != {}'.format(noise.shape, (self.D,)))
        nat.check(nat.lib.lmc_op_set_kernels(
            self._h, Q, nat.host_ptr(kinds), nat.host_ptr(params), nat.host_ptr(Bs), nat.host_ptr(noise)))
        self.Q = Q
        self.kernels_on_device = True
        self.kernel_param_counts = [2 if d[0] == KERNEL_KINDS['StdPeriodic'] else 1 for d in desc]
        self._set_factors(Q, coreg_vecs, coreg_diags)

    def kernel_tops(self, deriv=False):
        """The tops the device evaluated for set_kernels, [Q, m] (or [sum p_q, m] derivatives)."""
        cnt = nat.lib.lmc_op_num_kernel_tops(self._h, int(deriv))
        out = np.empty((cnt, self.m))
        nat.check(nat.lib.lmc_op_kernel_tops(self._h, int(deriv), nat.host_ptr(out)))
        return out

    def grid_dists(self):
        """||z - z_0|| of the operator's grid, as the device evaluates it (interpolated_llgp.py:431)."""
        ax = [np.arange(m) * d for m, d in zip(self.grid_sizes, self.grid_deltas)]
        if self.ndim == 1:
            return ax[0]
        return np.sqrt(ax[0][:, None] ** 2 + ax[1][None, :] ** 2)

    def _set_factors(self, Q, coreg_vecs, coreg_diags):
        if coreg_vecs is not None and coreg_diags is not None:
            vecs = [np.atleast_2d(np.asarray(a, dtype=np.float64)) for a in coreg_vecs]
            ranks = nat.as_i32([len(a) for a in vecs])
            A = nat.as_f64(np.vstack(vecs))
            kappa = nat.as_f64(np.array(coreg_diags))
            if len(vecs) != Q or A.shape[1] != self.D or kappa.shape != (Q, self.D):
                raise ValueError('coregionalisation factors have the wrong shape')
            nat.check(nat.lib.lmc_op_set_coreg_factors(
                self._h, nat.host_ptr(ranks), nat.host_ptr(A), nat.host_ptr(kappa)))

    def perm(self):
        p = np.empty(self.n, dtype=np.int32)
        nat.check(nat.lib.lmc_op_perm(self._h, nat.host_ptr(p)))
        return p

    # ---- products --------------------------------------------------------
    def _block(self, V):
        V = nat.as_f64(V)
        if V.ndim == 1:
            V = V.reshape(1, -1)
        if V.ndim != 2 or V.shape[1] != self.n:
            raise ValueError('expected block of shape (P, {}), got {}'.format(self.n, V.shape))
        return V

    def mvm(self, V):
        """K~ applied to the rows of V ([P, n] or [n]); host in, host out."""
        single = np.ndim(V) == 1
        V = self._block(V)
        out = np.empty_like(V)
        nat.check(nat.lib.lmc_mvm_host(self._h, nat.host_ptr(V), self.n,
                                        V.shape[0], nat.host_ptr(out)))
        return out[0] if single else out

    def mvm_into(self, V, out):
        """Host-buffer product without allocations: V, out are C-contiguous float64
        [P, n] arrays (pin them, e.g. torch's pin_memory().numpy(), for full PCIe
        overlap).  Copy-in, product and copy-out are pipelined over column chunks."""
        assert V.flags['C_CONTIGUOUS'] and out.flags['C_CONTIGUOUS'] and V.shape == out.shape
        nat.check(nat.lib.lmc_mvm_host(self._h, nat.host_ptr(V), self.n, V.shape[0], nat.host_ptr(out)))
        return out

    def matvec(self, x):
        return self.mvm(np.asarray(x, dtype=np.float64).reshape(-1))

    def matmat(self, X):
        """K~ X for an [n, P] array, the reference's Matrix.matmat (linalg/matrix.py:27-41).  A C-ordered
        array goes to the device as it is (point-major entry point, no host transposition); a
        Fortran-ordered one is P contiguous columns and takes the pipelined column path."""
        X = np.asarray(X, dtype=np.float64)
        if X.ndim != 2 or X.shape[0] != self.n:
            raise ValueError('expected block of shape ({}, P), got {}'.format(self.n, X.shape))
        if X.flags['F_CONTIGUOUS'] and not X.flags['C_CONTIGUOUS']:
            return self.mvm(X.T).T
        X = np.ascontiguousarray(X)
        out = np.empty_like(X)
        P = X.shape[1]
        nat.check(nat.lib.lmc_mvm_rows_host(self._h, nat.host_ptr(X), P, P, nat.host_ptr(out), P))
        return out

    def matmat_device(self, X, out=None):
        """X: torch float64 CUDA tensor [n, P], point-major: unit stride along the columns, any row stride
        >= P (a column slice of a wider block is fine).  Stream ordered."""
        torch = nat.require_cuda()
        assert X.is_cuda and X.dtype == torch.float64 and X.dim() == 2 and X.shape[0] == self.n
        P = X.shape[1]
        assert (P == 1 or X.stride(1) == 1) and X.stride(0) >= P
        if out is None:
            out = torch.empty((self.n, P), dtype=torch.float64, device=X.device)
        assert (P == 1 or out.stride(1) == 1) and out.stride(0) >= P and out.shape == X.shape
        nat.check(nat.lib.lmc_mvm_rows(self._h, _dev(X), X.stride(0), P, _dev(out), out.stride(0),
                                        nat.current_stream_ptr()))
        return out

    def mvm_device(self, V, out=None):
        """V: torch float64 CUDA tensor [P, n] (row stride = n). Stream ordered."""
        torch = nat.require_cuda()
        assert V.is_cuda and V.dtype == torch.float64 and V.is_contiguous()
        if out is None:
            out = torch.empty_like(V)
        nat.check(nat.lib.lmc_mvm(self._h, _dev(V), V.shape[1], V.shape[0],
                                   _dev(out), nat.current_stream_ptr()))
        return out

    def mvm_sorted_device(self, V, out=None):
        """Same as mvm_device with V and the result in the operator's sorted point order
        (column i is the caller's point perm()[i]) -- the solver's native layout."""
        torch = nat.require_cuda()
        assert V.is_cuda and V.dtype == torch.float64 and V.is_contiguous()
        if out is None:
            out = torch.empty_like(V)
        nat.check(nat.lib.lmc_mvm_sorted(self._h, _dev(V), V.shape[1], V.shape[0],
                                          _dev(out), nat.current_stream_ptr()))
        return out

    def to_grid_device(self, V):
        torch = nat.require_cuda()
        G = torch.empty((V.shape[0], self.D * self.m), dtype=torch.float64, device=V.device)
        nat.check(nat.lib.lmc_to_grid(self._h, _dev(V), V.shape[1], V.shape[0],
                                       _dev(G), nat.current_stream_ptr()))
        return G

    def grid_mvm_device(self, G):
        torch = nat.require_cuda()
        out = torch.empty_like(G)
        nat.check(nat.lib.lmc_grid_mvm(self._h, _dev(G), G.shape[0], _dev(out),
                                        nat.current_stream_ptr()))
        return out

    def from_grid_device(self, G):
        torch = nat.require_cuda()
        out = torch.empty((G.shape[0], self.n), dtype=torch.float64, device=G.device)
        nat.check(nat.lib.lmc_from_grid(self._h, _dev(G), G.shape[0], _dev(out),
                                         self.n, nat.current_stream_ptr()))
        return out

    # ---- solves ----------------------------------------------------------
    def diagonal(self):
        """diag(K~) in the caller's point order (exact, without forming K~; lmc_op_diagonal)."""
        out = np.empty(self.n)
        nat.check(nat.lib.lmc_op_diagonal(self._h, nat.host_ptr(out)))
        return out

    def minres(self, RHS, tol=1e-4, maxiter=None, check_every=100, precond=None):
        """Batched Iterative.solve (approx/iterative.py:24-62) on the rows of
        RHS.  Returns (X, iters, resid, istop).  precond='jacobi': scipy's M = diag(K~)^-1."""
        RHS = self._block(RHS)
        P = RHS.shape[0]
        if precond is not None:
            torch = nat.require_cuda()
            X, iters, resid, istop = self.minres_device(torch.as_tensor(RHS, device='cuda'), tol=tol, maxiter=maxiter,
                                                        check_every=check_every, precond=precond)
            return X.cpu().numpy(), iters, resid, istop
        X = np.empty_like(RHS)
        iters = np.zeros(P, dtype=np.int32)
        resid = np.zeros(P, dtype=np.float64)
        istop = np.zeros(P, dtype=np.int32)
        nat.check(nat.lib.lmc_minres_host(
            self._h, nat.host_ptr(RHS), self.n, P, nat.host_ptr(X), float(tol),
            int(self.n if maxiter is None else maxiter), int(check_every),
            nat.host_ptr(iters), nat.host_ptr(resid), nat.host_ptr(istop)))
        return X, iters, resid, istop

    def cg(self, RHS, tol=1e-4, maxiter=None, check_every=100):
        """Batched Iterative.solve(..., minres=False) (scipy's cg behind the reference wrapper) on the
        rows of RHS.  Returns (X, iters, resid, info)."""
        RHS = self._block(RHS)
        P = RHS.shape[0]
        X = np.empty_like(RHS)
        iters = np.zeros(P, dtype=np.int32)
        resid = np.zeros(P, dtype=np.float64)
        info = np.zeros(P, dtype=np.int32)
        nat.check(nat.lib.lmc_cg_host(
            self._h, nat.host_ptr(RHS), self.n, P, nat.host_ptr(X), float(tol),
            int(self.n if maxiter is None else maxiter), int(check_every),
            nat.host_ptr(iters), nat.host_ptr(resid), nat.host_ptr(info)))
        return X, iters, resid, info

    PRECONDITIONERS = {None: 0, 'none': 0, 'jacobi': 1}     # LMC_PRECOND_* of include/lmc_b200.h

    def minres_lanczos_device(self, RHS, k, tol=1e-4, maxiter=None, check_every=100):
        """minres_device that also returns the Lanczos tridiagonals of the solves (lmc_minres_lanczos):
        (X, iters, resid, istop, tridiag [P, k, 2], beta1 [P])."""
        torch = nat.require_cuda()
        assert RHS.is_cuda and RHS.dtype == torch.float64 and RHS.is_contiguous()
        P = RHS.shape[0]
        X = torch.empty_like(RHS)
        iters = np.zeros(P, dtype=np.int32)
        resid = np.zeros(P, dtype=np.float64)
        istop = np.zeros(P, dtype=np.int32)
        tri = np.zeros((P, int(k), 2), dtype=np.float64)
        beta1 = np.zeros(P, dtype=np.float64)
        nat.check(nat.lib.lmc_minres_lanczos(
            self._h, _dev(RHS), RHS.shape[1], P, _dev(X), float(tol),
            int(self.n if maxiter is None else maxiter), int(check_every),
            nat.host_ptr(iters), nat.host_ptr(resid), nat.host_ptr(istop), int(k), nat.host_ptr(tri),
            nat.host_ptr(beta1), nat.current_stream_ptr()))
        return X, iters, resid, istop, tri, beta1

    def minres_device(self, RHS, tol=1e-4, maxiter=None, check_every=100, precond=None):
        torch = nat.require_cuda()
        assert RHS.is_cuda and RHS.dtype == torch.float64 and RHS.is_contiguous()
        if precond not in self.PRECONDITIONERS:
            raise ValueError('unknown preconditioner {!r}'.format(precond))
        P = RHS.shape[0]
        X = torch.empty_like(RHS)
        iters = np.zeros(P, dtype=np.int32)
        resid = np.zeros(P, dtype=np.float64)
        istop = np.zeros(P, dtype=np.int32)
        nat.check(nat.lib.lmc_minres_pre(
            self._h, _dev(RHS), RHS.shape[1], P, _dev(X), float(tol),
            int(self.n if maxiter is None else maxiter), int(check_every), self.PRECONDITIONERS[precond],
            nat.host_ptr(iters), nat.host_ptr(resid), nat.host_ptr(istop),
            nat.current_stream_ptr()))
        return X, iters, resid, istop

    # ---- gradient contractions --------------------------------------------
    def grad_grams_device(self, alpha, R, RINV, extra_tops=()):
        """alpha [n], R/RINV [N, n] CUDA tensors; extra_tops: derivative tops, or None for the
        derivative tops of the kernels given to set_kernels (evaluated on the device).
        Returns (quad[T,D,D], trace[T,D,D], nquad[D], ntrace[D]) numpy, T = Q + number of extra tops."""
        torch = nat.require_cuda()
        N = 0 if R is None else R.shape[0]
        on_device = extra_tops is None
        if on_device:
            if not self.kernels_on_device:
                raise ValueError('derivative tops on the device need set_kernels')
            ext = np.zeros((nat.lib.lmc_op_num_kernel_tops(self._h, 1), 0))
        else:
            ext = nat.as_f64(np.array([np.asarray(t, dtype=np.float64).ravel() for t in extra_tops])
                             if len(extra_tops) else np.zeros((0, self.m)))
        T = self.Q + ext.shape[0]
        D = self.D
        quad = np.zeros((T, D, D))
        trace = np.zeros((T, D, D))
        nquad = np.zeros(D)
        ntrace = np.zeros(D)
        assert alpha.is_contiguous() and alpha.dtype == torch.float64
        if N:
            assert R.is_contiguous() and RINV.is_contiguous() and R.shape == RINV.shape
        if on_device:
            nat.check(nat.lib.lmc_grad_grams_kernels(
                self._h, _dev(alpha), _dev(R) if N else None, _dev(RINV) if N else None, self.n, N,
                nat.host_ptr(quad), nat.host_ptr(trace), nat.host_ptr(nquad), nat.host_ptr(ntrace),
                nat.current_stream_ptr()))
            return quad, trace, nquad, ntrace
        nat.check(nat.lib.lmc_grad_grams(
            self._h, _dev(alpha), _dev(R) if N else None, _dev(RINV) if N else None,
            self.n, N, ext.shape[0], nat.host_ptr(ext) if ext.shape[0] else None,
            nat.host_ptr(quad), nat.host_ptr(trace), nat.host_ptr(nquad), nat.host_ptr(ntrace),
            nat.current_stream_ptr()))
        return quad, trace, nquad, ntrace

    def grad_grams(self, alpha, R, RINV, extra_tops=()):
        torch = nat.require_cuda()
        dev = torch.device('cuda')
        a = torch.as_tensor(nat.as_f64(alpha), device=dev)
        Rt = torch.as_tensor(nat.as_f64(R), device=dev) if len(R) else None
        Ri = torch.as_tensor(nat.as_f64(RINV), device=dev) if len(R) else None
        return self.grad_grams_device(a, Rt, Ri, extra_tops)


def assemble_gradients(coreg_vecs, coreg_mats, kernel_param_counts, N, quad, trace, nquad, ntrace):
    """Chain rule of ApproxLMCLikelihood's gradient families
    (runlmc/lmc/likelihood.py:48-96) from the Gram matrices:
    dL/dtheta = 0.5 (<C, quad_t> - <C, trace_t> / N)  (derivative.py:5-6,
    stochastic_deriv.py:69-78).

    Returns (coreg_vec_grads, coreg_diag_grads, kernel_grads, noise_grad)."""
    Q = len(coreg_vecs)
    D = len(nquad)
    Nn = max(N, 1)
    M = [0.5 * (quad[t] - trace[t] / Nn) for t in range(quad.shape[0])]
    cv, cd, kg = [], [], []
    for q, a in enumerate(coreg_vecs):
        a = np.atleast_2d(a)
        g = np.zeros(a.shape)
        for i, ai in enumerate(a):
            # dA = e_j ai^T + ai e_j^T  ->  <dA, M> = M[j,:].ai + ai.M[:,j]
            g[i] = M[q].dot(ai) + ai.dot(M[q])
        cv.append(g)
        cd.append(np.diag(M[q]).copy())                 # C = E_ii
    t = Q
    for q in range(Q):
        gq = []
        for _ in range(kernel_param_counts[q]):
            gq.append(float(np.sum(coreg_mats[q] * M[t])))   # C = B_q, top = dk_q/dtheta
            t += 1
        kg.append(gq)
    noise = 0.5 * (np.asarray(nquad) - np.asarray(ntrace) / Nn)
    return cv, cd, kg, noise
