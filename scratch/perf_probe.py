import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from runlmc_b200 import synthetic, _native as nat
from runlmc_b200.fused import FusedLMC
def probe(name, P, cpl, iters=5, **kw):
    t0 = time.time()
    prob = synthetic.make_problem(name, seed=1234, cells_per_lengthscale=cpl, **kw)
    t1 = time.time()
    op = FusedLMC(prob.Xs, prob.grids); op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    t2 = time.time()
    print(f'{name}: n={prob.n} gen {t1-t0:.1f}s create {t2-t1:.2f}s max_tile_pts', nat.lib.lmc_op_max_tile_points(op._h))
    V = torch.randn(P, prob.n, dtype=torch.float64, device='cuda')
    out = torch.empty_like(V)
    for _ in range(2): op.mvm_device(V, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): op.mvm_device(V, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bytes_alg = 16.0 * prob.n * P + 8 * prob.ndim * prob.n
    print(f'  block MVM P={P}: {ms:.3f} ms  -> {P/ms*1e3:.0f} MVM/s, alg GB/s {bytes_alg/ms/1e6:.0f}')
    nat.profile_begin()
    for _ in range(iters): op.mvm_device(V, out)
    prof = nat.profile_end()
    tot = sum(v[0] for v in prof.values())
    for k, (m, c) in prof.items(): print(f'    {k:18s} {m/iters:8.3f} ms/step  {c//iters:4d} launches  {100*m/tot:5.1f}%')
    return op, prob
op, prob = probe('D', 65, 8)
op, prob = probe('E', 129, 6)
# minres timing on E (few iterations)
R = torch.tensor(np.vstack([prob.y[None], prob.probes[:16]]), device='cuda')
torch.cuda.synchronize(); t = time.time()
X, it, res, st = op.minres_device(R, tol=1e-4, maxiter=20, check_every=100)
torch.cuda.synchronize(); dt = time.time() - t
print('E minres 17 rhs x 20 iters: %.3f s -> %.0f iter*rhs/s' % (dt, 17*20/dt), it[:4], res[:4])
# same E problem with inputs pre-sorted into the operator's order (perm = identity)
perm = op.perm()
off = np.concatenate([[0], np.cumsum(prob.lens)])
Xall = np.vstack(prob.Xs)[perm]
prob.Xs = [Xall[off[d]:off[d+1]] for d in range(prob.D)]
op2 = FusedLMC(prob.Xs, prob.grids); op2.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
assert np.all(op2.perm() == np.arange(prob.n))
V = torch.randn(129, prob.n, dtype=torch.float64, device='cuda'); out = torch.empty_like(V)
for _ in range(2): op2.mvm_device(V, out)
nat.profile_begin()
for _ in range(5): op2.mvm_device(V, out)
prof = nat.profile_end()
print('E with identity perm:')
for k, (m, c) in prof.items(): print(f'    {k:18s} {m/5:8.3f} ms/step')
