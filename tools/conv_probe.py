#!/usr/bin/env python
"""How many MINRES iterations the reference's stopping rules need on a workload as a function of the
kernel lengthscale (grid cells) and the noise level eps (noise ~ 1 / Gamma(1 + 1/eps, 1), the reference
benchmark's parameter, benchmarks/benchlib/bench.py:111-115).  Used to pick bench.py's converging
gradient row.

    python tools/conv_probe.py E 1.5,4 0.1,1 [jacobi|-] [noise scales]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from runlmc_b200 import kern, synthetic  # noqa: E402
from runlmc_b200.fused import FusedLMC  # noqa: E402


def main():
    wl = sys.argv[1]
    cpls = [float(x) for x in sys.argv[2].split(',')]
    epss = [float(x) for x in sys.argv[3].split(',')]
    pre = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != '-' else None
    scales = [float(x) for x in sys.argv[5].split(',')] if len(sys.argv) > 5 else [1.0]
    op = None
    for cpl in cpls:
        for eps, scale in [(e, s) for e in epss for s in scales]:
            prob = synthetic.make_problem(wl, seed=1234, cells_per_lengthscale=cpl, eps=eps, N=4, noise_scale=scale)
            if op is None:
                op = FusedLMC(prob.Xs, prob.grids)
            op.set_kernels([kern.RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs,
                           prob.coreg_diags)
            R = torch.as_tensor(np.vstack([prob.y[None], prob.probes]), device='cuda')
            for p in ([None, pre] if pre else [None]):
                t = time.time()
                X, it, res, istop = op.minres_device(R, tol=1e-4, maxiter=4000, precond=p)
                torch.cuda.synchronize()
                print('%s cpl=%g eps=%g noise=[%.3g..%.3g] precond=%s: iters %s resid %s istop %s  %.2fs' % (
                    wl, cpl, eps, prob.noise.min(), prob.noise.max(), p, it.tolist(),
                    ['%.1e' % r for r in res], istop.tolist(), time.time() - t), flush=True)


if __name__ == '__main__':
    main()
