import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from test_fused_gpu import fused_from_problem
from test_oracle_golden import GOLDEN_PROBLEMS, oracle_operator
from conftest import rel_err, load_golden
from oracle import lmc_oracle as orc
for name in ['lmc_A', 'lmc_2d', 'lmc_B']:
    g = load_golden(name); prob = GOLDEN_PROBLEMS[name](); op = fused_from_problem(prob)
    RHS = np.vstack([prob.y[None, :], g['probes']])
    X, iters, resid, istop = op.minres(RHS, tol=1e-4)
    want = np.vstack([g['alpha'][None, :], g['inv_probes']])
    _, ref = oracle_operator(prob)
    print(name, 'ref ctr', int(g['solve_y_ctr']), 'ref err', float(g['solve_y_err']))
    for i, (x, w) in enumerate(zip(X, want)):
        rr = np.linalg.norm(RHS[i] - ref.matvec(w))
        print('  col', i, 'iters', iters[i], 'istop', istop[i], 'resid %.3e' % resid[i], 'ref resid %.3e' % rr, 'relerr %.2e' % rel_err(x, w))
