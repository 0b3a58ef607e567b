#!/usr/bin/env python
"""Run exactly one pass of the hot path inside a cudaProfilerStart/Stop window, for
`ncu --profile-from-start off` captures (see tools/ncu_step.sh).

    python tools/one_step.py E mvm            # block product, caller's order
    python tools/one_step.py E mvm_rows       # block product, caller's order, point-major [n, P] block
    python tools/one_step.py E mvm_sorted     # block product, operator's sorted order
    python tools/one_step.py D minres 2       # K solver iterations
    python tools/one_step.py E grad           # gradient contractions
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from runlmc_b200 import synthetic  # noqa: E402
from runlmc_b200.fused import FusedLMC  # noqa: E402

CPL = {'A': 4, 'B': 8, 'C': 5, 'D': 2, 'E': 1.5}


def main():
    wl, mode = sys.argv[1], sys.argv[2]
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    P = int(os.environ.get('LMC_P', '0'))
    prob = synthetic.make_problem(wl, seed=1234, cells_per_lengthscale=CPL[wl])
    op = FusedLMC(prob.Xs, prob.grids)
    op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    V = torch.as_tensor(np.vstack([prob.y[None], prob.probes]), device='cuda')
    if P:
        V = V[:P].contiguous()
    out = torch.empty_like(V)
    if mode == 'mvm':
        fn = lambda: op.mvm_device(V, out)
    elif mode == 'mvm_rows':
        X = V.t().contiguous()
        Y = torch.empty_like(X)
        fn = lambda: op.matmat_device(X, Y)
    elif mode == 'mvm_sorted':
        fn = lambda: op.mvm_sorted_device(V, out)
    elif mode == 'minres':
        fn = lambda: op.minres_device(V, tol=1e-4, maxiter=k, check_every=100)
    elif mode == 'grad':
        fn = lambda: op.grad_grams_device(V[0], V[1:], out[1:], prob.top_grads_flat()
                                          if hasattr(prob, 'top_grads_flat') else ())
    else:
        raise SystemExit('unknown mode ' + mode)
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == '__main__':
    main()
