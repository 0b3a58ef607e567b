import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from runlmc_b200 import synthetic, _native as nat
from runlmc_b200.fused import FusedLMC
def presort(prob):
    op = FusedLMC(prob.Xs, prob.grids); perm = op.perm()
    off = np.concatenate([[0], np.cumsum(prob.lens)])
    Xall = np.vstack(prob.Xs)[perm]
    prob.Xs = [Xall[off[d]:off[d+1]] for d in range(prob.D)]
    return prob
def probe(name, P, cpl, iters=5, sort=False):
    prob = synthetic.make_problem(name, seed=1234, cells_per_lengthscale=cpl)
    if sort: prob = presort(prob)
    op = FusedLMC(prob.Xs, prob.grids); op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    V = torch.randn(P, prob.n, dtype=torch.float64, device='cuda'); out = torch.empty_like(V)
    for _ in range(2): op.mvm_device(V, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): op.mvm_device(V, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bytes_alg = 16.0 * prob.n * P + 8 * prob.ndim * prob.n
    print(f'{name} sorted={sort} P={P}: {ms:.3f} ms -> {P/ms*1e3:.0f} MVM/s, alg {bytes_alg/ms/1e6:.0f} GB/s ({bytes_alg/ms/1e6/65.4:.1f}% of 6540)')
    nat.profile_begin()
    for _ in range(iters): op.mvm_device(V, out)
    prof = nat.profile_end()
    print('   ', ' '.join(f'{k}={m/iters:.3f}' for k, (m, c) in prof.items()))
    return op, prob
for name, P, cpl in (('D', 65, 8), ('E', 129, 6), ('B', 17, 8), ('C', 17, 5)):
    for sort in (False, True):
        op, prob = probe(name, P, cpl, sort=sort)
    R = torch.tensor(np.vstack([prob.y[None], prob.probes[:P-1]]), device='cuda')
    torch.cuda.synchronize(); t = time.time()
    X, it, res, st = op.minres_device(R, tol=1e-4, maxiter=50, check_every=100)
    torch.cuda.synchronize(); dt = time.time() - t
    nat.profile_begin()
    X, it, res, st = op.minres_device(R, tol=1e-4, maxiter=50, check_every=100)
    prof = nat.profile_end()
    print(f'  minres {P} rhs x 50 it: {dt*1e3:.1f} ms -> {P*50/dt:.0f} iter*rhs/s;', ' '.join(f'{k}={m/50:.3f}' for k, (m, c) in prof.items()))
