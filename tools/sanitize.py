#!/usr/bin/env python
"""One small pass over every hot-path entry point (products in both orders, the three stages, MINRES,
CG, gradient Gram stage, host-buffer product) on the small stand-ins of configs A / D / E, for
compute-sanitizer (tools/sanitize.sh).  Sizes are tiny: racecheck slows kernels by ~100x."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from runlmc_b200 import kern, synthetic  # noqa: E402
from runlmc_b200.fused import FusedLMC  # noqa: E402

CASES = [('A', dict(cells_per_lengthscale=4)),
         ('d_small', dict(cells_per_lengthscale=6)),
         ('e_small', dict(cells_per_lengthscale=3)),
         ('e_small', dict(cells_per_lengthscale=3, lens=[700, 0, 650], grid=[40, 24], edge=True)),
         # 512 x 512 embedding: register transforms, bulk-copy (TMA) row passes, the odd-column scatter (7 columns)
         ('e_small', dict(cells_per_lengthscale=4, lens=[300, 0, 280], grid=[130, 200], edge=True))]


def main():
    only = sys.argv[1:] or None
    for name, kw in CASES:
        if only and name not in only:
            continue
        prob = synthetic.make_problem(name, seed=3, **kw)
        op = FusedLMC(prob.Xs, prob.grids)
        op.set_kernels([kern.RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs,
                       prob.coreg_diags)
        Vh = np.vstack([prob.y[None], prob.probes])
        V = torch.as_tensor(Vh, device='cuda')
        KV = op.mvm_device(V)
        KR = op.matmat_device(V.t().contiguous())      # point-major block: row staging, row-writing gathers
        assert torch.equal(KR.t(), KV)
        op.matmat(np.ascontiguousarray(Vh[:5].T))
        op.mvm_sorted_device(V)
        G = op.to_grid_device(V[:3].contiguous())
        op.grid_mvm_device(G)
        op.from_grid_device(G)
        op.mvm(Vh[:3])
        X, it, res, _ = op.minres_device(V, tol=1e-4, maxiter=12, check_every=5)
        op.minres_device(V, tol=1e-4, maxiter=6, check_every=5, precond='jacobi')
        op.minres_lanczos_device(V, 8, tol=1e-4, maxiter=6, check_every=5)
        op.cg(Vh[:3], tol=1e-4, maxiter=6)
        op.grad_grams_device(X[0], V[1:], X[1:], None)
        torch.cuda.synchronize()
        print(name, 'ok', float(KV.abs().sum()), it[:3])


if __name__ == '__main__':
    main()
