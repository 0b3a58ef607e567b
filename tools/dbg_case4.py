import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from runlmc_b200 import kern, synthetic
from runlmc_b200.fused import FusedLMC
from oracle import lmc_oracle as orc
def rel(a,b): return float(np.linalg.norm(a-b)/np.linalg.norm(b))
prob = synthetic.make_problem('e_small', seed=3, cells_per_lengthscale=3, lens=[700, 0, 650], grid=[40, 24], edge=True)
op = FusedLMC(prob.Xs, prob.grids)
op.set_kernels([kern.RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
spec = orc.KernelSpec(['rbf'] * prob.Q, [[g] for g in prob.gammas], prob.coreg_vecs, prob.coreg_diags, prob.noise)
ref = orc.build_operator(spec, prob.Xs, prob.grids, rep='sum')
Vh = np.vstack([prob.y[None], prob.probes]); V = torch.as_tensor(Vh, device='cuda')
for rep in range(3):
    KV = op.mvm_device(V).cpu().numpy()
    print('mvm', [ '%.1e' % rel(KV[c], ref.matvec(Vh[c])) for c in range(len(Vh))])
    KVs = op.mvm_sorted_device(V[:, torch.as_tensor(op.perm().astype(np.int64), device='cuda')].contiguous()).cpu().numpy()
    print('mvm_sorted', ['%.1e' % rel(KVs[c], ref.matvec(Vh[c])[op.perm()]) for c in range(len(Vh))])
    for P in (1,2,3,7):
        G = op.to_grid_device(V[:P].contiguous()).cpu().numpy()
        print(' to_grid P=%d'%P, ['%.1e' % rel(G[c], ref.WT.dot(Vh[c])) for c in range(P)])
        Gd = torch.as_tensor(np.array([ref.WT.dot(Vh[c]) for c in range(P)]), device='cuda')
        KG = op.grid_mvm_device(Gd).cpu().numpy()
        print(' grid_mvm', ['%.1e' % rel(KG[c], ref.grid_matvec(Gd[c].cpu().numpy())) for c in range(P)])
        F = op.from_grid_device(Gd).cpu().numpy()
        print(' from_grid', ['%.1e' % rel(F[c], ref.W.dot(Gd[c].cpu().numpy())) for c in range(P)])
