#!/usr/bin/env python
"""Developer timing of the block product and the solver, per kernel family, in the caller's
order and in the operator's sorted order.  Not the bench contract (see bench.py).

    python tools/devbench.py E [P] [--minres K] [--cpl X]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from runlmc_b200 import synthetic, _native as nat  # noqa: E402
from runlmc_b200.fused import FusedLMC  # noqa: E402

CPL = {'A': 4, 'B': 8, 'C': 5, 'D': 2, 'E': 1.5}


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nat.profile_begin()
    for _ in range(iters):
        fn()
    prof = nat.profile_end()
    return ms, {k: (m / iters, c // iters) for k, (m, c) in prof.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('workload')
    ap.add_argument('P', nargs='?', type=int, default=0)
    ap.add_argument('--minres', type=int, default=0, help='time K MINRES iterations')
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--cpl', type=float, default=0)
    args = ap.parse_args()
    prob = synthetic.make_problem(args.workload, seed=1234, cells_per_lengthscale=args.cpl or CPL[args.workload])
    P = args.P or prob.N + 1
    op = FusedLMC(prob.Xs, prob.grids)
    op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    V = torch.randn(P, prob.n, dtype=torch.float64, device='cuda')
    out = torch.empty_like(V)
    alg = 16.0 * prob.n * P + 8.0 * prob.ndim * prob.n
    VR = V.t().contiguous()
    outR = torch.empty_like(VR)
    for label, fn in (('caller order', lambda: op.mvm_device(V, out)),
                      ('caller order, point-major', lambda: op.matmat_device(VR, outR)),
                      ('sorted order', lambda: op.mvm_sorted_device(V, out))):
        ms, prof = timed(fn, args.iters)
        print('%s %s P=%d: %.3f ms/step  %.0f MVM*RHS/s  alg %.0f GB/s' % (
            args.workload, label, P, ms, P / ms * 1e3, alg / ms / 1e6))
        print('    ' + '  '.join('%s=%.3f(%d)' % (k, m, c) for k, (m, c) in prof.items()))
    if args.minres:
        R = torch.tensor(np.vstack([prob.y[None], prob.probes[:P - 1]]), device='cuda')
        op.minres_device(R, tol=1e-4, maxiter=3, check_every=100)
        torch.cuda.synchronize()
        t = time.time()
        op.minres_device(R, tol=1e-4, maxiter=args.minres, check_every=100)
        torch.cuda.synchronize()
        dt = time.time() - t
        nat.profile_begin()
        op.minres_device(R, tol=1e-4, maxiter=args.minres, check_every=100)
        prof = nat.profile_end()
        print('  minres %d rhs x %d it: %.1f ms -> %.3f ms/it, %.0f iter*rhs/s' % (
            P, args.minres, dt * 1e3, dt * 1e3 / args.minres, P * args.minres / dt))
        print('    ' + '  '.join('%s=%.3f' % (k, m / args.minres) for k, (m, c) in prof.items()))


if __name__ == '__main__':
    main()
