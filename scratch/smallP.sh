for P in 17 33 65; do timeout 200 python tools/devbench.py E $P --minres 20 | grep -v caller -A0 ; done
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0,'.')
from runlmc_b200 import synthetic
from runlmc_b200.fused import FusedLMC
import torch; torch.cuda.init(); torch.zeros(1,device='cuda')
prob = synthetic.make_problem('E', seed=1234, cells_per_lengthscale=1.5)
for k in range(2):
    t=time.time(); op = FusedLMC(prob.Xs, prob.grids); torch.cuda.synchronize(); print('E op create #%d %.3f s'%(k,time.time()-t))
PY
