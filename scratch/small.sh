for w in A B C; do timeout 300 python tools/devbench.py $w --minres 200 --iters 20; done
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0,'.')
from runlmc_b200 import synthetic
from runlmc_b200.fused import FusedLMC
prob = synthetic.make_problem('E', seed=1234, cells_per_lengthscale=1.5)
import torch; torch.cuda.init()
t=time.time(); op = FusedLMC(prob.Xs, prob.grids); torch.cuda.synchronize(); print('E op create %.3f s'%(time.time()-t))
t=time.time(); op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags); torch.cuda.synchronize(); print('E set_params %.3f s'%(time.time()-t))
t=time.time(); op.set_params(prob.tops, prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags); torch.cuda.synchronize(); print('E set_params again %.3f s'%(time.time()-t))
PY
