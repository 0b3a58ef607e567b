"""Mirror of runlmc/lmc/grid_kernel.py: assemble the SKI-LMC operator.

`gen_grid_kernel` returns the same tree the reference builds
(SumMatrix([GridKernel..., Diag(noise)]), grid_kernel.py:49-74) -- every node is
a device-backed runlmc_b200.linalg class -- and, when the tree is the standard
SKI-LMC operator, attaches ONE fused CUDA operator per active-dimension group
(`_fused` for the usual single group: `matvec`, `Iterative.solve` and the
likelihood use it instead of walking the tree; `FusedGroupsSumMatrix` for
several groups).  sum / bt / slfm are the same matrix (except the
reference's slfm tree with an empty coreg or diag part, which is left to the
tree, see _try_fuse); the fused operator serves all three.  The handle is shared
by the operators built from one interpolant, each of which re-binds its own
hyper-parameters before use (_ParamState)."""
import numpy as np

from ..approx.ski import SKI
from ..linalg.diag import Diag
from ..linalg.block_diag import BlockDiag
from ..linalg.block_matrix import SymmSquareBlockMatrix
from ..linalg.matrix import Matrix
from ..linalg.composition import Composition
from ..linalg.bttb import BTTB
from ..linalg.identity import Identity
from ..linalg.kronecker import Kronecker
from ..linalg.numpy_matrix import NumpyMatrix
from ..linalg.sum_matrix import SumMatrix
from ..fused import FusedLMC, kernel_descriptor


class GridKernel(Matrix):
    """One active-dimension group of the SKI-LMC operator in the reference's representation `ktype`
    (grid_kernel.py:22-41).  `grid_K` / `ski` are assembled on first access from a snapshot of the
    hyper-parameters taken here: when the fused operator serves the products (the usual case) the
    O(Q m) host kernel evaluation and the tree of up to D^2 BTTB nodes are never built."""

    def __init__(self, functional_kernel, grid_dists, interpolant, interpolantT, ktype, active_dim):
        n = interpolant.shape[0]
        super().__init__(n, n)
        assert ktype in ('sum', 'bt', 'slfm'), ktype
        self.ktype = ktype
        self._interp = (interpolant, interpolantT)
        self._build = (_FrozenKernel(functional_kernel, active_dim, grid_dists), grid_dists, active_dim)
        self._grid_K = self._ski = None

    @property
    def grid_K(self):
        if self._grid_K is None:
            fk, grid_dists, active_dim = self._build
            grid_k = fk.eval_kernels_fixed_dim(grid_dists, active_dim)
            if self.ktype == 'sum':
                self._grid_K = _gen_sum_grid(fk, grid_k, active_dim)
            elif self.ktype == 'bt':
                self._grid_K = _gen_bt_grid(fk, grid_k, active_dim)
            else:
                self._grid_K = _gen_slfm_grid(fk, grid_k, self._interp[0].shape[1], active_dim)
        return self._grid_K

    @property
    def ski(self):
        if self._ski is None:
            self._ski = SKI(self.grid_K, *self._interp)
        return self._ski

    def _apply_dev(self, X):
        return self.ski._apply_dev(X)


class _FrozenKernel:
    """What a GridKernel needs of a FunctionalKernel for one active-dimension group, with the
    hyper-parameters copied at construction (the optimiser updates the FunctionalKernel in place,
    an operator built earlier must keep its own values like the reference's does)."""

    def __init__(self, fk, active_dim, grid_dists):
        import copy
        idxs = list(fk.active_dims[active_dim])
        self.D, self.Q = fk.D, fk.Q
        self.active_dims = {active_dim: idxs}
        self.coreg_vecs = [np.array(a, dtype=float) for a in fk.coreg_vecs]
        self.coreg_diags = [np.array(k, dtype=float) for k in fk.coreg_diags]
        self.num_lmc = {active_dim: fk.num_lmc[active_dim]}
        self.num_indep = {active_dim: fk.num_indep[active_dim]}
        self._non_indep = list(fk.filter_non_indep_idxs(idxs))
        kernels = getattr(fk, '_kernels', None)
        self._kernels = copy.deepcopy(kernels) if kernels is not None else None
        # duck-typed stand-ins without kernel objects cannot be snapshotted: evaluate them now
        self._grid_k = None if kernels is not None else fk.eval_kernels_fixed_dim(grid_dists, active_dim)

    def eval_kernels_fixed_dim(self, dists, active_dim):
        if self._kernels is not None:
            return np.array([self._kernels[k].from_dist(dists) for k in self.active_dims[active_dim]])
        return self._grid_k

    def coreg_mats(self, active_dim=None):
        idxs = range(len(self.coreg_vecs)) if active_dim is None else self.active_dims[active_dim]
        return [self.coreg_vecs[i].T.dot(self.coreg_vecs[i]) + np.diag(self.coreg_diags[i]) for i in idxs]

    def filter_non_indep_idxs(self, idxs):
        return [k for k in idxs if k in self._non_indep]


class _ParamState:
    """The hyper-parameters of ONE gen_grid_kernel call, as the fused handle takes them.  The handle
    (point sort, workspace) is shared by every operator built from the same interpolant; `bind` makes
    it hold THIS operator's parameters before it is used, so operators built earlier keep their own
    values (K(theta + h) and K(theta - h) side by side, a prediction operator held across optimiser
    steps) -- like the reference, whose gen_grid_kernel returns independent operators."""

    def __init__(self, descs, tops, Bs, noise, coreg_vecs, coreg_diags):
        self.descs = descs
        self.tops = None if tops is None else [np.array(t, dtype=float) for t in tops]
        self.Bs = [np.array(B, dtype=float) for B in Bs]
        self.noise = np.array(noise, dtype=float)
        self.coreg_vecs = [np.array(a, dtype=float) for a in coreg_vecs]
        self.coreg_diags = [np.array(k, dtype=float) for k in coreg_diags]

    def bind(self, fused):
        if getattr(fused, '_bound_state', None) is not self:
            if self.descs is not None:
                fused.set_kernel_descriptors(self.descs, self.Bs, self.noise, self.coreg_vecs, self.coreg_diags)
            else:
                fused.set_params(self.tops, self.Bs, self.noise, self.coreg_vecs, self.coreg_diags)
            fused._bound_state = self
        return fused


class FusedSumMatrix(SumMatrix):
    """SumMatrix([GridKernel, Diag(noise)]) collapsed into one device handle."""

    def __init__(self, Ks, fused, state=None):
        super().__init__(Ks)
        self._handle = fused
        self._state = state

    @property
    def _fused(self):
        """The device operator holding this matrix's hyper-parameters (None after unpickling)."""
        if self._handle is None:
            return None
        return self._state.bind(self._handle) if self._state is not None else self._handle

    def _apply_dev(self, X):
        fused = self._fused
        if fused is None:           # unpickled: walk the (device-backed) tree like any SumMatrix
            return super()._apply_dev(X)
        return fused.mvm_device(X.contiguous())

    def matmat(self, X):
        """[n, P] block as the caller holds it: a C-ordered array goes through the point-major entry point
        (lmc_mvm_rows_host), without the two host transpositions of the generic Matrix.matmat."""
        fused = self._fused
        if fused is None:
            return super().matmat(X)
        X = np.asarray(X)
        if X.ndim != 2 or X.shape[0] != self.shape[1]:
            raise ValueError('dimension mismatch: {} vs {}'.format(X.shape, self.shape))
        return fused.matmat(X)

    def __getstate__(self):
        state = super().__getstate__()
        state['_handle'] = None      # device handles are not picklable
        state['_state'] = None
        return state


class FusedGroupsSumMatrix(SumMatrix):
    """SumMatrix([GridKernel_g ..., Diag(noise)]) for kernels on several active-dimension groups
    (reference grid_kernel.py:49-74 loops over fk.active_dims): one fused device operator per group --
    W_g (sum_{q in g} B_q (x) T_q) W_g^T, the first one carrying the noise term -- instead of one launch per
    tree node.  Solves go through the block solver's callback path with this product."""

    def __init__(self, Ks, handles, states):
        super().__init__(Ks)
        self._handles = handles
        self._states = states

    def _apply_dev(self, X):
        if self._handles is None:       # unpickled
            return super()._apply_dev(X)
        X = X.contiguous()
        total = None
        for handle, state in zip(self._handles, self._states):
            Y = state.bind(handle).mvm_device(X)
            if total is None:
                total = Y
            else:
                total += Y
        return total

    def __getstate__(self):
        state = super().__getstate__()
        state['_handles'] = None
        state['_states'] = None
        return state


def representation(fk, active_dim):
    """The reference's selection rule (grid_kernel.py:52-64)."""
    if fk.Q == 1:
        return 'sum'
    tot_rank = fk.total_rank(active_dim)
    corr = fk.D if (not fk.num_lmc[active_dim] and not fk.num_indep[active_dim]) else 0
    return 'slfm' if tot_rank + fk.D < fk.D ** 2 + corr else 'bt'


def gen_grid_kernel(fk, grid_dists, interpolants, lens_per_output):
    grid_kerns = {}
    for active_dim in fk.active_dims.keys():
        interpolant, interpolantT = interpolants[active_dim]
        grid_kerns[active_dim] = GridKernel(fk, grid_dists[active_dim], interpolant, interpolantT,
                                            representation(fk, active_dim), active_dim)
    noise = Diag(np.repeat(fk.noise, lens_per_output))
    ls = list(grid_kerns.values())
    ls.append(noise)
    groups = list(fk.active_dims.keys())
    fused = [_try_fuse(fk, grid_dists, interpolants, ad, noise_on=(i == 0)) for i, ad in enumerate(groups)]
    if all(f is not None for f in fused):
        if len(fused) == 1:
            handle, state = fused[0]
            state.bind(handle)
            return FusedSumMatrix(ls, handle, state), grid_kerns
        return FusedGroupsSumMatrix(ls, [h for h, _ in fused], [st for _, st in fused]), grid_kerns
    return SumMatrix(ls), grid_kerns


def _try_fuse(fk, grid_dists, interpolants, active_dim, noise_on=True):
    """(handle, parameter state) of the fused operator of one active-dimension group of 1 or 2 input
    dimensions whose interpolant carries its geometry, else None.  noise_on: this group's operator carries the
    noise term (exactly one group of a model does)."""
    W = interpolants[active_dim][0]
    geom = getattr(W, 'lmc_geometry', None)
    if geom is None or len(geom[1]) not in (1, 2) or fk.D > 16:
        return None
    idxs = fk.active_dims[active_dim]
    if representation(fk, active_dim) == 'slfm' and (
            not fk.filter_non_indep_idxs(idxs) or
            (fk.num_lmc[active_dim] == 0 and fk.num_indep[active_dim] == 0)):
        # the reference's slfm tree puts Identity(m) in place of an empty coreg or diag part
        # (grid_kernel.py:84-86, 101-103): that operator is W (sum_q B_q x T_q + I) W^T + noise, which
        # the fused handle does not represent -- leave it to the tree, which reproduces it
        return None
    Xs, grids = geom
    cache = getattr(W, '_lmc_fused', None)
    if cache is None:
        cache = FusedLMC(Xs, grids)          # X-dependent sort happens once per model
        W._lmc_fused = cache
    coreg_vecs = [fk.coreg_vecs[i] for i in idxs]
    coreg_diags = [fk.coreg_diags[i] for i in idxs]
    noise = fk.noise if noise_on else np.zeros(fk.D)
    kerns = _device_kernels(fk, idxs, cache, grid_dists[active_dim])
    if kerns is not None:
        # per-step setup on the device: kernel values are evaluated where the spectra are computed
        state = _ParamState([kernel_descriptor(k) for k in kerns], None, fk.coreg_mats(active_dim), noise,
                            coreg_vecs, coreg_diags)
    else:
        grid_k = fk.eval_kernels_fixed_dim(grid_dists[active_dim], active_dim)
        state = _ParamState(None, list(grid_k), fk.coreg_mats(active_dim), noise, coreg_vecs, coreg_diags)
    return cache, state


def _device_kernels(fk, idxs, fused, dists):
    """The kernels of this group if the device can evaluate all of them on its own grid distances
    (they must be the distances the caller passed), else None."""
    kernels = getattr(fk, '_kernels', None)
    if kernels is None:
        return None
    kerns = [kernels[i] for i in idxs]
    if any(kernel_descriptor(k) is None for k in kerns):
        return None
    dists = np.asarray(dists)
    # the comparison is O(m) on the host: remember the verdict for this very array (models pass the same
    # `dists` object on every optimiser step)
    seen = getattr(fused, '_dists_checked', None)
    if seen is None or seen[0] is not dists:
        own = fused.grid_dists()
        ok = dists.shape == own.shape and np.allclose(dists, own, rtol=1e-13,
                                                      atol=1e-13 * max(1.0, float(own.max())))
        fused._dists_checked = (dists, bool(ok))
        seen = fused._dists_checked
    return kerns if seen[1] else None


# ---- the three representations of the grid operator K_UU (reference grid_kernel.py:77-136) ----------------
# grid_k: [Q_group, *grid shape] kernel values of the group's kernels, in the order of fk.active_dims[active_dim].

def _bttb_of(values):
    return BTTB(np.ravel(values), values.shape)


def _gen_sum_grid(fk, grid_k, active_dim):
    """sum_q B_q (x) T_q, one Kronecker term per kernel."""
    return SumMatrix([Kronecker(NumpyMatrix(B), _bttb_of(k)) for B, k in zip(fk.coreg_mats(active_dim), grid_k)])


def _gen_bt_grid(fk, grid_k, active_dim):
    """D x D blocks, block (i, j) the BTTB of sum_q B_q[i, j] k_q (symmetric: built once per unordered pair)."""
    D = fk.D
    grid_shape = grid_k.shape[1:]
    mixed = np.einsum('qij,qm->ijm', np.array(fk.coreg_mats(active_dim)), grid_k.reshape(len(grid_k), -1))
    upper = {(i, j): BTTB(mixed[i, j], grid_shape) for i in range(D) for j in range(i, D)}
    return SymmSquareBlockMatrix([[upper[min(i, j), max(i, j)] for j in range(D)] for i in range(D)])


def _gen_slfm_grid(fk, grid_k, m, active_dim):
    """(A* (x) I) blockdiag(T_q per coregionalisation vector) (A*^T (x) I)  +  blockdiag_d(sum_q kappa_q[d] T_q);
    an empty part is Identity(m), as in the reference (see _try_fuse)."""
    return SumMatrix([_gen_coreg_Ks(fk, grid_k, m, active_dim), _gen_diag_Ks(fk, grid_k, m, active_dim)])


def _gen_coreg_Ks(fk, grid_k, m, active_dim):
    vecs = [fk.coreg_vecs[q] for q in fk.filter_non_indep_idxs(fk.active_dims[active_dim])]
    if len(vecs) == 0:
        return Identity(m)
    # the non-independent kernels come first in the group, so vecs[i] belongs to grid_k[i]
    per_vector = [t for a, k in zip(vecs, grid_k) for t in [_bttb_of(k)] * len(a)]
    stacked = np.vstack(vecs)                                # [sum R_q, D]
    eye = Identity(int(np.prod(grid_k.shape[1:])))
    return Composition([Kronecker(NumpyMatrix(stacked.T), eye), BlockDiag(per_vector),
                        Kronecker(NumpyMatrix(stacked), eye)])


def _gen_diag_Ks(fk, grid_k, m, active_dim):
    if not (fk.num_lmc[active_dim] or fk.num_indep[active_dim]):
        return Identity(m)
    kappa = np.array([fk.coreg_diags[q] for q in fk.active_dims[active_dim]])      # [Q_group, D]
    per_output = kappa.T.dot(grid_k.reshape(len(grid_k), -1))                       # [D, m]
    return BlockDiag([BTTB(top, grid_k.shape[1:]) for top in per_output])
