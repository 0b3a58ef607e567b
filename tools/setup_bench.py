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


def reference_call_sequence(name='E'):
    """The reference's own set-up sequence (interpolated_llgp.py:431-437, 192-207) through the mirror:
    multi_interpolant -> transpose().tocsr() -> gen_grid_kernel, first call (point sort on the device) and a later
    call with new hyper-parameters (the per-step cost).  The CSR is lazy: nothing here assembles its 4^d n entries."""
    from runlmc_b200.approx.interpolation import multi_interpolant
    from runlmc_b200.lmc.functional_kernel import FunctionalKernel
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    prob = synthetic.make_problem(name, seed=1234, cells_per_lengthscale={'E': 1.5, 'D': 2}.get(name, 4))
    fk = FunctionalKernel(D=prob.D, lmc_kernels=[kern.RBF(g) for g in prob.gammas], lmc_ranks=[1] * prob.Q)
    fk.noise = prob.noise
    fk.coreg_vecs = prob.coreg_vecs
    fk.coreg_diags = prob.coreg_diags
    fk.set_input_dim(prob.ndim)
    ad = tuple(range(prob.ndim))
    torch.cuda.synchronize()
    t = time.perf_counter()
    W = multi_interpolant(prob.Xs, *prob.grids)
    WT = W.transpose().tocsr()
    t_w = time.perf_counter() - t
    t = time.perf_counter()
    K, _ = gen_grid_kernel(fk, {ad: prob.dists}, {ad: (W, WT)}, prob.lens)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t
    fk.noise = 1.1 * prob.noise
    t = time.perf_counter()
    K2, _ = gen_grid_kernel(fk, {ad: prob.dists}, {ad: (W, WT)}, prob.lens)
    torch.cuda.synchronize()
    t_step = time.perf_counter() - t
    v = np.ones(prob.n)
    K2.matvec(v)
    print('%s reference call sequence: multi_interpolant + transpose().tocsr() %.1f ms (CSR assembled: %s), '
          'gen_grid_kernel first call %.1f ms, per step %.2f ms' % (
              name, t_w * 1e3, W.materialized, t_first * 1e3, t_step * 1e3))
    t = time.perf_counter()
    _ = W.nnz
    print('%s   (assembling the CSR when something does read it: %.2f s, %d nonzeros)' % (
        name, time.perf_counter() - t, W.nnz))


def reference_gradient_step(name='E'):
    """One optimiser step the way the reference's model takes it (interpolated_llgp.py:192-207): gen_grid_kernel,
    ApproxLMCLikelihood(..., StochasticDerivService) -- probes drawn inside generate() -- and the four gradient
    families, on the well-conditioned hyper-parameters of bench.py's converging gradient row."""
    from runlmc_b200.approx.interpolation import multi_interpolant
    from runlmc_b200.lmc.functional_kernel import FunctionalKernel
    from runlmc_b200.lmc.grid_kernel import gen_grid_kernel
    from runlmc_b200.lmc.likelihood import ApproxLMCLikelihood
    from runlmc_b200.lmc.stochastic_deriv import StochasticDerivService
    from runlmc_b200.util.inline_pool import InlinePool
    kw = {'E': dict(cells_per_lengthscale=1.5, eps=1.0, noise_scale=300.0),
          'D': dict(cells_per_lengthscale=2, eps=1.0, noise_scale=100.0)}.get(name, dict(cells_per_lengthscale=4))
    prob = synthetic.make_problem(name, seed=1234, **kw)
    fk = FunctionalKernel(D=prob.D, lmc_kernels=[kern.RBF(g) for g in prob.gammas], lmc_ranks=[1] * prob.Q)
    fk.noise = prob.noise
    fk.coreg_vecs = prob.coreg_vecs
    fk.coreg_diags = prob.coreg_diags
    fk.set_input_dim(prob.ndim)
    ad = tuple(range(prob.ndim))
    W = multi_interpolant(prob.Xs, *prob.grids)
    WT = W.transpose().tocsr()
    for label, dev_probes in (('probes from numpy (reference stream)', False), ('probes drawn on the device', True),
                              ('probes from numpy (reference stream)', False), ('probes drawn on the device', True)):
        np.random.seed(1)
        torch.manual_seed(1)
        torch.cuda.synchronize()
        t = time.perf_counter()
        K, _ = gen_grid_kernel(fk, {ad: prob.dists}, {ad: (W, WT)}, prob.lens)
        svc = StochasticDerivService(None, InlinePool(None), prob.N, 1e-4, device_probes=dev_probes)
        lik = ApproxLMCLikelihood(fk, K, {ad: prob.dists}, {ad: (W, WT)}, prob.Ys, svc)
        grads = fk.update_gradient(lik)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        print('%s one optimiser step through the reference call sequence, %s: %.3f s (noise gradient[0] = %.6g)'
              % (name, label, dt, grads['noise'][0]))
        del lik, K, svc


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'E'
    main()
    reference_call_sequence(which)
    reference_gradient_step(which)
