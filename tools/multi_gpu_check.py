#!/usr/bin/env python
"""N-rank gradient == 1-rank gradient on real GPUs (run under torchrun, one rank per GPU):

  (a) Gram stage: the trace / quadratic Gram matrices from IDENTICAL solves, sharded over the ranks and
      all-reduced, against the same stage on one rank -- differs only in summation order;
  (b) the whole sharded evaluation (each rank solves its own probes: a column's complex-pair partner
      changes with the sharding, which moves a converged MINRES solution at the 1e-6 level, DESIGN.md
      section 3) against the 1-rank evaluation.

Rank 0 prints one JSON line.  Used by tests/test_multi_gpu.py and for the committed log in profiles/."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from runlmc_b200 import kern, synthetic  # noqa: E402
from runlmc_b200.distributed import allreduce_trace, shard_bounds, sharded_gradient  # noqa: E402
from runlmc_b200.fused import FusedLMC, assemble_gradients  # noqa: E402


def flat(grads):
    return np.concatenate([np.ravel(g) for g in grads[0]] + [np.ravel(g) for g in grads[1]] +
                          [np.ravel(g) for g in grads[2]] + [np.ravel(grads[3])])


def main():
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    name = sys.argv[1] if len(sys.argv) > 1 else 'medium'
    if name == 'medium':
        prob = synthetic.make_problem('e_small', seed=21, cells_per_lengthscale=1.5, eps=10.0, D=4,
                                      lens=[6000, 5000, 5500, 6500], grid=[64, 48], Q=3, N=22)
    else:
        prob = synthetic.make_problem(name, seed=1234, cells_per_lengthscale=1.0, eps=10.0, N=16)
    op = FusedLMC(prob.Xs, prob.grids)
    op.set_kernels([kern.RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs,
                   prob.coreg_diags)
    dev = torch.device('cuda', local)
    # ---- one rank (before the process group exists: no collective) ----
    g1, s1 = sharded_gradient(op, prob.y, prob.probes, None, prob.coreg_vecs, prob.coreg_mats(), tol=1e-4)
    RHS = torch.as_tensor(np.vstack([prob.y[None], prob.probes]), device=dev)
    X, _, _, _ = op.minres_device(RHS, tol=1e-4)
    quad1, trace1, nquad1, ntrace1 = op.grad_grams_device(X[0], RHS[1:], X[1:], None)
    dist.init_process_group('nccl', device_id=dev)
    # ---- (a) Gram stage on identical solves, sharded ----
    lo, hi = shard_bounds(prob.N, rank, world)
    quad, trace, nquad, ntrace = op.grad_grams_device(
        X[0], RHS[1 + lo:1 + hi].contiguous() if hi > lo else None, X[1 + lo:1 + hi].contiguous() if hi > lo else None, None)
    trace, ntrace, _, _ = allreduce_trace(trace, ntrace, 0.0, 0.0)
    ga = flat(assemble_gradients(prob.coreg_vecs, prob.coreg_mats(), op.kernel_param_counts, prob.N, quad, trace,
                                 nquad, ntrace))
    gref = flat(assemble_gradients(prob.coreg_vecs, prob.coreg_mats(), op.kernel_param_counts, prob.N, quad1,
                                   trace1, nquad1, ntrace1))
    err_gram = float(np.linalg.norm(ga - gref) / np.linalg.norm(gref))
    # ---- (b) the whole evaluation, probes sharded ----
    gN, sN = sharded_gradient(op, prob.y, prob.probes, None, prob.coreg_vecs, prob.coreg_mats(), tol=1e-4,
                              rank=rank, world=world)
    f1, fN = flat(g1), flat(gN)
    err_full = float(np.linalg.norm(fN - f1) / np.linalg.norm(f1))
    both = torch.tensor([err_gram, err_full], dtype=torch.float64, device=dev)
    dist.all_reduce(both, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({'world': world, 'problem': name, 'n': prob.n, 'probes': prob.N,
                          'gram_stage_rel_err': float(both[0]), 'full_gradient_rel_err': float(both[1]),
                          'mean_iterations_1rank': s1['iterations'], 'mean_iterations_Nrank': sN['iterations'],
                          'mean_residual_1rank': s1['solv_error'], 'mean_residual_Nrank': sN['solv_error'],
                          'gradient_l2': float(np.linalg.norm(f1))}))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
