import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from runlmc_b200 import synthetic
from runlmc_b200.fused import FusedLMC
name = sys.argv[1]; cpl = float(sys.argv[2]); nrhs = int(sys.argv[3])
prob = synthetic.make_problem(name, seed=1234, cells_per_lengthscale=cpl)
op = FusedLMC(prob.Xs, prob.grids); op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
R = torch.tensor(np.vstack([prob.y[None], prob.probes[:nrhs-1]]), device='cuda')
torch.cuda.synchronize(); t = time.time()
X, it, res, st = op.minres_device(R, tol=1e-4, maxiter=3000)
torch.cuda.synchronize(); dt = time.time() - t
print(name, 'cpl', cpl, 'noise', prob.noise.round(3), 'iters', it, 'istop', st, 'resid', res.round(6), '%.2fs' % dt, '%.0f iter*rhs/s' % (it.sum()/dt))
