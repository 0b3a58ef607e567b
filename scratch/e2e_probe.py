"""e2e (host buffers) rate of the block product for the current LMC_HOST_CHUNK_MB."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from runlmc_b200 import synthetic
from runlmc_b200.fused import FusedLMC
name = sys.argv[1] if len(sys.argv) > 1 else 'E'
prob = synthetic.make_problem(name, seed=1234, cells_per_lengthscale={'E': 1.5, 'D': 2}[name])
op = FusedLMC(prob.Xs, prob.grids)
op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
Vh = np.ascontiguousarray(np.vstack([prob.y[None, :], prob.probes]))
Vp = torch.as_tensor(Vh).pin_memory(); Op = torch.empty_like(Vp).pin_memory()
Vn, On = Vp.numpy(), Op.numpy()
op.mvm_into(Vn, On); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    op.mvm_into(Vn, On)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 5
print('%s chunk_mb=%s: %.2f ms/step, %.0f MVM*RHS/s, %.1f GB/s each way' % (
    name, os.environ.get('LMC_HOST_CHUNK_MB', '64'), dt * 1e3, Vh.shape[0] / dt, Vh.nbytes / dt / 1e9))
