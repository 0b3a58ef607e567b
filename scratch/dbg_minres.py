import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from test_fused_gpu import PROBLEMS, fused_from_problem, oracle_from_problem
from conftest import rel_err
from oracle import lmc_oracle as orc
for name in ['A', 'd_small']:
    prob = PROBLEMS[name](); op = fused_from_problem(prob); _, ref = oracle_from_problem(prob)
    RHS = np.vstack([prob.y[None, :], prob.probes[:3]])
    for P in (1, 2, 4):
        for k in (1, 2, 3, 7, 25):
            X, iters, resid, istop = op.minres(RHS[:P], tol=1e-4, maxiter=k, check_every=10**6)
            errs = []
            for b, x in zip(RHS[:P], X):
                xr, _, itn, _ = orc.minres(ref.matvec, b, 1e-10, k)
                errs.append(rel_err(x, xr))
            print(name, 'P', P, 'k', k, 'iters', iters, 'istop', istop, 'err', ['%.1e' % e for e in errs])
