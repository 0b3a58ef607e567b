#!/usr/bin/env python
"""Wall time and per-family device time of the gradient Gram stage (lmc_grad_grams_kernels).

    python tools/gram_bench.py E [N]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from runlmc_b200 import kern, synthetic, _native as nat  # noqa: E402
from runlmc_b200.fused import FusedLMC  # noqa: E402

CPL = {'A': 4, 'B': 8, 'C': 5, 'D': 2, 'E': 1.5}


def main():
    wl = sys.argv[1]
    prob = synthetic.make_problem(wl, seed=1234, cells_per_lengthscale=CPL[wl])
    N = int(sys.argv[2]) if len(sys.argv) > 2 else prob.N
    op = FusedLMC(prob.Xs, prob.grids)
    op.set_kernels([kern.RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs,
                   prob.coreg_diags)
    R = torch.as_tensor(prob.probes[:N], device='cuda')
    X = torch.randn(N + 1, prob.n, dtype=torch.float64, device='cuda')
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        op.grad_grams_device(X[0], R, X[1:], None)
        torch.cuda.synchronize()
        print('%s N=%d gram stage wall %.1f ms' % (wl, N, 1e3 * (time.perf_counter() - t0)), flush=True)
    nat.profile_begin()
    op.grad_grams_device(X[0], R, X[1:], None)
    prof = nat.profile_end()
    print('   ' + '  '.join('%s=%.2f(%d)' % (k, m, c) for k, (m, c) in prof.items()),
          ' sum %.1f ms' % sum(m for m, c in prof.values()))


if __name__ == '__main__':
    main()
