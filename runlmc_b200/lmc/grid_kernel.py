"""Mirror of runlmc/lmc/grid_kernel.py: assemble the SKI-LMC operator.

`gen_grid_kernel` returns the same tree the reference builds
(SumMatrix([GridKernel..., Diag(noise)]), grid_kernel.py:49-74) -- every node is
a device-backed runlmc_b200.linalg class -- and, when the tree is the standard
single-active-dimension-group SKI-LMC operator, attaches ONE fused CUDA
operator (`_fused`) that `matvec`, `Iterative.solve` and the likelihood use
instead of walking the tree.  sum / bt / slfm are the same matrix; the fused
operator serves all three."""
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
    def __init__(self, functional_kernel, grid_dists, interpolant, interpolantT, ktype, active_dim):
        n = interpolant.shape[0]
        super().__init__(n, n)
        grid_k = functional_kernel.eval_kernels_fixed_dim(grid_dists, active_dim)
        if ktype == 'sum':
            self.grid_K = _gen_sum_grid(functional_kernel, grid_k, active_dim)
        elif ktype == 'bt':
            self.grid_K = _gen_bt_grid(functional_kernel, grid_k, active_dim)
        elif ktype == 'slfm':
            self.grid_K = _gen_slfm_grid(functional_kernel, grid_k, interpolant.shape[1], active_dim)
        else:
            assert False, ktype
        self.ktype = ktype
        self.ski = SKI(self.grid_K, interpolant, interpolantT)

    def _apply_dev(self, X):
        return self.ski._apply_dev(X)


class FusedSumMatrix(SumMatrix):
    """SumMatrix([GridKernel, Diag(noise)]) collapsed into one device handle."""

    def __init__(self, Ks, fused):
        super().__init__(Ks)
        self._fused = fused

    def _apply_dev(self, X):
        return self._fused.mvm_device(X.contiguous())

    def __getstate__(self):
        state = super().__getstate__()
        state['_fused'] = None      # device handles are not picklable
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
    fused = _try_fuse(fk, grid_dists, interpolants, lens_per_output)
    if fused is not None:
        return FusedSumMatrix(ls, fused), grid_kerns
    return SumMatrix(ls), grid_kerns


def _try_fuse(fk, grid_dists, interpolants, lens_per_output):
    """One fused operator when there is a single active-dimension group of 1 or 2
    input dimensions and the interpolant carries its geometry."""
    if len(fk.active_dims) != 1:
        return None
    (active_dim,) = fk.active_dims.keys()
    W = interpolants[active_dim][0]
    geom = getattr(W, 'lmc_geometry', None)
    if geom is None or len(geom[1]) not in (1, 2) or fk.D > 16:
        return None
    Xs, grids = geom
    cache = getattr(W, '_lmc_fused', None)
    if cache is None:
        cache = FusedLMC(Xs, grids)          # X-dependent sort happens once per model
        W._lmc_fused = cache
    idxs = fk.active_dims[active_dim]
    factors = dict(coreg_vecs=[fk.coreg_vecs[i] for i in idxs],
                   coreg_diags=[fk.coreg_diags[i] for i in idxs])
    kerns = _device_kernels(fk, idxs, cache, grid_dists[active_dim])
    if kerns is not None:
        # per-step setup on the device: kernel values are evaluated where the spectra are computed
        cache.set_kernels(kerns, fk.coreg_mats(active_dim), fk.noise, **factors)
    else:
        grid_k = fk.eval_kernels_fixed_dim(grid_dists[active_dim], active_dim)
        cache.set_params(list(grid_k), fk.coreg_mats(active_dim), fk.noise, **factors)
    return cache


def _device_kernels(fk, idxs, fused, dists):
    """The kernels of this group if the device can evaluate all of them on its own grid distances
    (they must be the distances the caller passed), else None."""
    kernels = getattr(fk, '_kernels', None)
    if kernels is None:
        return None
    kerns = [kernels[i] for i in idxs]
    if any(kernel_descriptor(k) is None for k in kerns):
        return None
    own = fused.grid_dists()
    dists = np.asarray(dists)
    if dists.shape != own.shape or not np.allclose(dists, own, rtol=1e-13, atol=1e-13 * max(1.0, float(own.max()))):
        return None
    return kerns


def _gen_slfm_grid(fk, grid_k, m, active_dim):
    return SumMatrix([_gen_coreg_Ks(fk, grid_k, m, active_dim), _gen_diag_Ks(fk, grid_k, m, active_dim)])


def _gen_coreg_Ks(fk, grid_k, m, active_dim):
    kidxs = fk.active_dims[active_dim]
    all_coreg = [fk.coreg_vecs[i] for i in fk.filter_non_indep_idxs(kidxs)]
    if not all_coreg:
        return Identity(m)
    ranks = [len(c) for c in all_coreg]
    A_star = np.vstack(all_coreg).T
    I_m = Identity(int(np.prod(grid_k.shape[1:])))
    left = Kronecker(NumpyMatrix(A_star), I_m)
    right = Kronecker(NumpyMatrix(A_star.T), I_m)
    toeps = []
    for top, r in zip(grid_k[:len(all_coreg)], ranks):
        t = BTTB(top.ravel(), top.shape)
        toeps.extend([t] * r)
    return Composition([left, BlockDiag(toeps), right])


def _gen_diag_Ks(fk, grid_k, m, active_dim):
    if fk.num_lmc[active_dim] == 0 and fk.num_indep[active_dim] == 0:
        return Identity(m)
    kidxs = fk.active_dims[active_dim]
    diags = np.column_stack([fk.coreg_diags[k] for k in kidxs])
    Q = grid_k.shape[0]
    diag_tops = diags.dot(grid_k.reshape(Q, -1))
    return BlockDiag([BTTB(top, grid_k.shape[1:]) for top in diag_tops])


def _gen_bt_grid(fk, grid_k, active_dim):
    Bs = np.array(fk.coreg_mats(active_dim))
    Q = grid_k.shape[0]
    bt = np.tensordot(Bs, grid_k.reshape(Q, -1), axes=(0, 0))
    sizes = grid_k.shape[1:]
    D = fk.D
    blocks = [[None] * D for _ in range(D)]
    for i in range(D):
        for j in range(i, D):
            blocks[i][j] = blocks[j][i] = BTTB(bt[i, j], sizes)
    return SymmSquareBlockMatrix(blocks)


def _gen_sum_grid(fk, grid_k, active_dim):
    Q = grid_k.shape[0]
    tops = grid_k.reshape(Q, -1)
    sizes = grid_k.shape[1:]
    return SumMatrix([Kronecker(NumpyMatrix(A), BTTB(top, sizes))
                      for A, top in zip(fk.coreg_mats(active_dim), tops)])
