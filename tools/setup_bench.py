#!/usr/bin/env python
"""Times the per-model and per-step operator setup, host path vs device path.

    python tools/setup_bench.py [E|D]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from runlmc_b200 import synthetic, kern  # noqa: E402
from runlmc_b200.fused import FusedLMC  # noqa: E402


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best, out


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'E'
    prob = synthetic.make_problem(name, seed=1234, cells_per_lengthscale={'E': 1.5, 'D': 2}.get(name, 4))
    torch.cuda.init()
    FusedLMC(prob.Xs, prob.grids)   # warm-up: context, module load
    th, op_h = timed(lambda: FusedLMC(prob.Xs, prob.grids, build='host'))
    td, op_d = timed(lambda: FusedLMC(prob.Xs, prob.grids, build='device'))
    print('%s op create: host sort %.1f ms, device sort %.1f ms (includes the %.1f MB upload of X)'
          % (name, th * 1e3, td * 1e3, prob.n * prob.ndim * 8 / 1e6))
    assert np.array_equal(op_h.perm(), op_d.perm())
    kerns = [kern.RBF(g) for g in prob.gammas]
    Bs = prob.coreg_mats()

    def host_step():
        tops = [k.from_dist(prob.dists) for k in kerns]
        op_h.set_params(tops, Bs, prob.noise, prob.coreg_vecs, prob.coreg_diags)

    tp, _ = timed(host_step, 5)
    tk, _ = timed(lambda: op_d.set_kernels(kerns, Bs, prob.noise, prob.coreg_vecs, prob.coreg_diags), 5)
    print('%s per-step setup: host tops + upload %.2f ms, device tops %.2f ms' % (name, tp * 1e3, tk * 1e3))


if __name__ == '__main__':
    main()
