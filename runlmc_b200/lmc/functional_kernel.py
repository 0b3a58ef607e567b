"""A paramz-free container for the hyper-parameters of an LMC kernel
    K(x_i, x_j)[d, d'] = sum_q B_q[d, d'] k_q(|x_i - x_j|),   B_q = A_q^T A_q + diag(kappa_q),
exposing the attributes the hot path reads from the reference's FunctionalKernel
(runlmc/lmc/functional_kernel.py: D, Q, noise, coreg_vecs, coreg_diags, coreg_mats, eval_kernels*,
eval_kernel_gradients, active_dims, num_lmc / num_slfm / num_indep, total_rank, get_active_dims,
filter_non_indep_idxs, update_gradient).  Parameters are plain numpy arrays; priors, transforms and the
optimiser plumbing of paramz are out of scope."""
import numpy as np
import scipy.stats

_LMC, _SLFM, _INDEP = 'lmc', 'slfm', 'indep'


class FunctionalKernel:
    """Three kinds of terms, stored in this order:

    * `lmc_kernels` with `lmc_ranks`: A_q is [rank, D] (random start), kappa_q = 1;
    * `slfm_kernels`: rank one, kappa_q = 0;
    * `indep_gp` with `indep_gp_index` (default 0, 1, ...): A_q = 0, kappa_q = e_index.

    :raises ValueError: as the reference does for a missing D, no kernels, mismatched or non-positive ranks,
        mismatched independent-GP indices."""

    def __init__(self, D=None, lmc_kernels=None, lmc_ranks=None, slfm_kernels=None,
                 indep_gp=None, indep_gp_index=None, name='kern'):
        self.name = name
        if not D:
            raise ValueError('D should be specified')
        groups = {_LMC: list(lmc_kernels or ()), _SLFM: list(slfm_kernels or ()), _INDEP: list(indep_gp or ())}
        if not any(groups.values()):
            raise ValueError('Number of kernels should be >0')
        ranks = list(lmc_ranks or ())
        if len(ranks) != len(groups[_LMC]):
            raise ValueError('# LMC kernels should equal # LMC ranks')
        if any(r <= 0 for r in ranks):
            raise ValueError('LMC ranks not positive')
        where = list(range(len(groups[_INDEP]))) if indep_gp_index is None else list(indep_gp_index)
        if len(where) != len(groups[_INDEP]):
            raise ValueError('indep GP number of kernels should match indices')
        self.D = D
        start = scipy.stats.truncnorm(-1, 1)
        self._kernels, self._kinds, self._coreg_vecs, self._coreg_diags = [], [], [], []
        for kern, rank in zip(groups[_LMC], ranks):
            self._add(kern, _LMC, start.rvs(size=(rank, D)), np.ones(D))
        for kern in groups[_SLFM]:
            self._add(kern, _SLFM, start.rvs(size=(1, D)), np.zeros(D))
        for kern, d in zip(groups[_INDEP], where):
            self._add(kern, _INDEP, np.zeros((1, D)), np.eye(D)[d])
        self._noise = np.full(D, 0.1)
        self.P = None
        self.gradient = None
        # filled by set_input_dim: kernel indices / counts per tuple of active input dimensions
        self.active_dims = {}
        self.num_lmc, self.num_slfm, self.num_indep = {}, {}, {}

    def _add(self, kern, kind, vecs, diag):
        self._kernels.append(kern)
        self._kinds.append(kind)
        self._coreg_vecs.append(np.array(vecs, dtype=float))
        self._coreg_diags.append(np.array(diag, dtype=float))

    def set_input_dim(self, P):
        """Fix the input dimension and group the kernels by their active dimensions (a kernel without
        `active_dims` uses all of them)."""
        if self.P == P:
            return
        if self.P is not None:
            raise ValueError('Cannot set input dimension twice')
        self.P = P
        counters = {_LMC: self.num_lmc, _SLFM: self.num_slfm, _INDEP: self.num_indep}
        for q, (kern, kind) in enumerate(zip(self._kernels, self._kinds)):
            dims = tuple(range(P)) if kern.active_dims is None else tuple(sorted(kern.active_dims))
            kern.active_dims = dims
            self.active_dims.setdefault(dims, []).append(q)
            for counter in counters.values():
                counter.setdefault(dims, 0)
            counters[kind][dims] += 1

    # ---- parameters ------------------------------------------------------
    @property
    def Q(self):
        return len(self._kernels)

    @property
    def noise(self):
        return self._noise

    @noise.setter
    def noise(self, value):
        self._noise[...] = value

    @property
    def coreg_vecs(self):
        return self._coreg_vecs

    @coreg_vecs.setter
    def coreg_vecs(self, values):
        for mine, theirs in zip(self._coreg_vecs, values):
            mine[...] = theirs

    @property
    def coreg_diags(self):
        return self._coreg_diags

    @coreg_diags.setter
    def coreg_diags(self, values):
        for mine, theirs in zip(self._coreg_diags, values):
            mine[...] = theirs

    def coreg_mats(self, active_dim=None):
        """B_q for every kernel, or for the kernels of one active-dimension group."""
        which = range(self.Q) if active_dim is None else self.active_dims[active_dim]
        return [self._coreg_vecs[q].T.dot(self._coreg_vecs[q]) + np.diag(self._coreg_diags[q]) for q in which]

    def total_rank(self, active_dim):
        assert self.P
        return sum(len(self._coreg_vecs[q]) for q in self.filter_non_indep_idxs(self.active_dims[active_dim]))

    def filter_non_indep_idxs(self, idxs):
        return [q for q in idxs if self._kinds[q] != _INDEP]

    def get_active_dims(self, q):
        return self._kernels[q].active_dims

    # ---- kernel values ---------------------------------------------------
    def eval_kernels(self, dists):
        """dists: {active dims: distances}; one array of kernel values per kernel."""
        assert self.P
        return [k.from_dist(dists[k.active_dims]) for k in self._kernels]

    def eval_kernels_fixed_dim(self, dists, active_dim):
        return np.array([self._kernels[q].from_dist(dists) for q in self.active_dims[active_dim]])

    def eval_kernel_gradients(self, dists):
        assert self.P
        return [k.kernel_gradient(dists[k.active_dims]) for k in self._kernels]

    def update_gradient(self, grads):
        """Collect the four gradient families of an LMCLikelihood (what the reference writes into its
        paramz parameters) and hand every kernel its own part."""
        assert self.P
        self.gradient = dict(coreg_vecs=grads.coreg_vec_gradients(), coreg_diags=grads.coreg_diags_gradients(),
                             kernels=grads.kernel_gradients(), noise=grads.noise_gradient())
        for kern, dk in zip(self._kernels, self.gradient['kernels']):
            kern.update_gradient(dk)
        return self.gradient
