#!/usr/bin/env python
"""Host-buffer product of config E through both layouts: P pinned columns (lmc_mvm_host) against one pinned
point-major [n, P] block (lmc_mvm_rows_host, strided 2-D copies).  Developer timing, not the bench contract."""
import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from runlmc_b200 import synthetic
from runlmc_b200.fused import FusedLMC
from runlmc_b200 import _native as nat
prob = synthetic.make_problem('E', seed=1234, cells_per_lengthscale=1.5)
op = FusedLMC(prob.Xs, prob.grids)
op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
Vh = np.vstack([prob.y[None], prob.probes])
P = Vh.shape[0]
Vp = torch.as_tensor(Vh).pin_memory(); Op = torch.empty_like(Vp).pin_memory()
Xp = torch.as_tensor(np.ascontiguousarray(Vh.T)).pin_memory(); Yp = torch.empty_like(Xp).pin_memory()
Vn, On, Xn, Yn = Vp.numpy(), Op.numpy(), Xp.numpy(), Yp.numpy()
def rows():
    nat.check(nat.lib.lmc_mvm_rows_host(op._h, nat.host_ptr(Xn), P, P, nat.host_ptr(Yn), P))
def cols():
    op.mvm_into(Vn, On)
for name, fn in (('cols', cols), ('rows', rows), ('cols', cols), ('rows', rows)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(name, '%.2f ms  %.0f MVM*RHS/s' % (dt * 1e3, P / dt))
print('max diff', np.abs(Yn.T - On).max())
