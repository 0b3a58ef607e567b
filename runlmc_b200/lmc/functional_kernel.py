"""Paramz-free FunctionalKernel with the surface the hot path consumes
(reference runlmc/lmc/functional_kernel.py:12-300: D, Q, noise, coreg_vecs,
coreg_diags, coreg_mats, eval_kernels*, eval_kernel_gradients, active_dims,
num_lmc/num_slfm/num_indep, total_rank, get_active_dims,
filter_non_indep_idxs, update_gradient).  Parameters are plain numpy arrays;
the optimiser/transform plumbing of paramz is out of scope."""
import numpy as np
import scipy.stats


class FunctionalKernel:
    _TRUNCNORM = scipy.stats.truncnorm(-1, 1)

    def __init__(self, D=None, lmc_kernels=None, lmc_ranks=None, slfm_kernels=None,
                 indep_gp=None, indep_gp_index=None, name='kern'):
        self.name = name
        if not D:
            raise ValueError('D should be specified')
        self.D = D
        if not lmc_kernels and not slfm_kernels and not indep_gp:
            raise ValueError('Number of kernels should be >0')
        lmc_kernels = list(lmc_kernels or [])
        lmc_ranks = list(lmc_ranks or [])
        if len(lmc_kernels) != len(lmc_ranks):
            raise ValueError('# LMC kernels should equal # LMC ranks')
        if not all(rank > 0 for rank in lmc_ranks):
            raise ValueError('LMC ranks not positive')
        slfm_kernels = list(slfm_kernels or [])
        indep_gp = list(indep_gp or [])
        indep_gp_index = list(indep_gp_index or range(len(indep_gp)))
        if len(indep_gp) != len(indep_gp_index):
            raise ValueError('indep GP number of kernels should match indices')
        self._kernels = lmc_kernels + slfm_kernels + indep_gp
        self._num_lmc = len(lmc_kernels)
        self._num_slfm = len(slfm_kernels)
        rnd = lambda r: FunctionalKernel._TRUNCNORM.rvs(size=(r, D))  # noqa: E731
        self._coreg_vecs = [rnd(r) for r in lmc_ranks] + [rnd(1) for _ in slfm_kernels] + \
            [np.zeros((1, D)) for _ in indep_gp]
        self._coreg_diags = [np.ones(D) for _ in lmc_kernels] + [np.zeros(D) for _ in slfm_kernels]
        for d in indep_gp_index:
            e = np.zeros(D)
            e[d] = 1
            self._coreg_diags.append(e)
        self._noise = 0.1 * np.ones(D)
        self.P = None
        self.active_dims = {}
        self.num_lmc, self.num_slfm, self.num_indep = {}, {}, {}
        self.gradient = None

    def set_input_dim(self, P):
        if self.P == P:
            return
        if self.P is not None:
            raise ValueError('Cannot set input dimension twice')
        self.P = P
        all_dims = tuple(range(P))
        for i, k in enumerate(self._kernels):
            k.active_dims = all_dims if k.active_dims is None else tuple(sorted(k.active_dims))
            self.active_dims.setdefault(k.active_dims, []).append(i)
            which = self.num_lmc if i < self._num_lmc else (
                self.num_slfm if i < self._num_lmc + self._num_slfm else self.num_indep)
            which[k.active_dims] = which.get(k.active_dims, 0) + 1
        for d in (self.num_lmc, self.num_slfm, self.num_indep):
            for ad in self.active_dims:
                d.setdefault(ad, 0)

    def update_gradient(self, grads):
        """Collect the gradients computed by an LMCLikelihood
        (functional_kernel.py:212-223)."""
        assert self.P
        self.gradient = {
            'coreg_vecs': grads.coreg_vec_gradients(),
            'coreg_diags': grads.coreg_diags_gradients(),
            'kernels': grads.kernel_gradients(),
            'noise': grads.noise_gradient()}
        for k, dk in zip(self._kernels, self.gradient['kernels']):
            k.update_gradient(dk)
        return self.gradient

    def total_rank(self, active_dim):
        assert self.P
        return sum(len(self._coreg_vecs[k]) for k in self.active_dims[active_dim]
                   if k < self._num_lmc + self._num_slfm)

    def eval_kernels(self, dists):
        assert self.P
        return [k.from_dist(dists[k.active_dims]) for k in self._kernels]

    def eval_kernels_fixed_dim(self, dists, active_dim):
        return np.array([self._kernels[k].from_dist(dists) for k in self.active_dims[active_dim]])

    def eval_kernel_gradients(self, dists):
        assert self.P
        return [k.kernel_gradient(dists[k.active_dims]) for k in self._kernels]

    @property
    def noise(self):
        return self._noise

    @noise.setter
    def noise(self, value):
        self._noise[:] = value

    @property
    def coreg_vecs(self):
        return self._coreg_vecs

    @coreg_vecs.setter
    def coreg_vecs(self, values):
        for cur, v in zip(self._coreg_vecs, values):
            cur[:] = v

    @property
    def coreg_diags(self):
        return self._coreg_diags

    @coreg_diags.setter
    def coreg_diags(self, values):
        for cur, v in zip(self._coreg_diags, values):
            cur[:] = v

    def coreg_mats(self, active_dim=None):
        cv, cd = self.coreg_vecs, self.coreg_diags
        if active_dim is not None:
            idxs = self.active_dims[active_dim]
            cv = [cv[i] for i in idxs]
            cd = [cd[i] for i in idxs]
        return [a.T.dot(a) + np.diag(k) for a, k in zip(cv, cd)]

    @property
    def Q(self):
        return len(self._kernels)

    def get_active_dims(self, q):
        return self._kernels[q].active_dims

    def filter_non_indep_idxs(self, idxs):
        lim = self._num_lmc + self._num_slfm
        return [k for k in idxs if k < lim]
