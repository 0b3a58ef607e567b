"""Probe sharding over the GPUs of one node.

Every rank holds the full operator and solves alpha = K^-1 y redundantly
(deterministic kernels => bitwise equal) plus its own contiguous, pair-aligned
slice of the N Hutchinson probes.  MINRES needs no communication; the only
collective is ONE all-reduce (sum, float64) of the per-hyper-parameter trace
partials [T*D*D + D] (+2 slots of solver statistics) per gradient evaluation --
NCCL over NVLink on GPUs, gloo in the CPU tests.  It replaces the reference's
pickled multiprocessing.Pool.starmap fan-out (stochastic_deriv.py:51-52)."""
import numpy as np


def shard_bounds(N, rank, world):
    """Contiguous [lo, hi) slice of N probes for `rank`; slice starts are even so
    complex RHS pairs (2j, 2j+1) never straddle ranks."""
    pairs = (N + 1) // 2
    base, rem = divmod(pairs, world)
    lo_p = rank * base + min(rank, rem)
    hi_p = lo_p + base + (1 if rank < rem else 0)
    return min(2 * lo_p, N), min(2 * hi_p, N)


def allreduce_trace(trace, ntrace, iters_sum, resid_sum, group=None):
    """Sum the probe-dependent partials over ranks.  Returns the reduced
    (trace, ntrace, iters_sum, resid_sum).  No-op when torch.distributed is not
    initialised (single process)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return trace, ntrace, iters_sum, resid_sum
    flat = np.concatenate([np.ravel(trace), np.ravel(ntrace), [iters_sum, resid_sum]])
    backend = dist.get_backend(group)
    device = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    t = torch.as_tensor(flat, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    flat = t.cpu().numpy()
    nt, nn = np.size(trace), np.size(ntrace)
    return (flat[:nt].reshape(np.shape(trace)), flat[nt:nt + nn].reshape(np.shape(ntrace)),
            float(flat[nt + nn]), float(flat[nt + nn + 1]))


def sharded_gradient(fused, y, probes, kernel_grad_tops, coreg_vecs, coreg_mats, tol=1e-4,
                     rank=0, world=1, group=None):
    """One stochastic gradient evaluation with the probes sharded over `world`
    ranks.  `probes` is the full [N, n] block (identical on every rank, e.g.
    from a shared seed); each rank touches only its slice.

    `kernel_grad_tops`: per kernel, the list of its derivative tops dk_q/dtheta on the grid -- or None
    when the operator evaluates its kernels on the device (FusedLMC.set_kernels).

    Returns (grads, stats): grads = (coreg_vec, coreg_diag, kernel, noise) as
    in ApproxLMCLikelihood, stats = dict(iterations, solv_error) (means over
    all N+1 solves, like Metrics, stochastic_deriv.py:42-45)."""
    from .fused import assemble_gradients
    N = len(probes)
    lo, hi = shard_bounds(N, rank, world)
    import torch
    local = np.asarray(probes[lo:hi], dtype=np.float64)
    # one upload of [y; local probes]; the solutions stay on the device for the Gram stage
    dev = torch.device('cuda', torch.cuda.current_device())
    RHS = torch.empty((1 + len(local), fused.n), dtype=torch.float64, device=dev)
    RHS[0].copy_(torch.as_tensor(np.ascontiguousarray(y, dtype=np.float64).reshape(-1)))
    if len(local):
        RHS[1:].copy_(torch.as_tensor(np.ascontiguousarray(local)))
    import time
    t0 = time.perf_counter()
    X, iters, resid, _ = fused.minres_device(RHS, tol=tol)      # returns after the solver's stream has drained
    t_solve = time.perf_counter() - t0
    if kernel_grad_tops is None:      # derivative tops of the kernels given to set_kernels, on the device
        extra, counts = None, fused.kernel_param_counts
    else:
        extra, counts = [t for ts in kernel_grad_tops for t in ts], [len(t) for t in kernel_grad_tops]
    t0 = time.perf_counter()
    quad, trace, nquad, ntrace = fused.grad_grams_device(
        X[0], RHS[1:] if len(local) else None, X[1:] if len(local) else None, extra)    # host results: synchronous
    t_gram = time.perf_counter() - t0
    alpha = X[0].cpu().numpy()
    it_sum = float(np.sum(iters[1:])) + (float(iters[0]) if rank == 0 else 0.0)
    rs_sum = float(np.sum(resid[1:])) + (float(resid[0]) if rank == 0 else 0.0)
    trace, ntrace, it_sum, rs_sum = allreduce_trace(trace, ntrace, it_sum, rs_sum, group)
    grads = assemble_gradients(coreg_vecs, coreg_mats, counts, N, quad, trace, nquad, ntrace)
    stats = {'iterations': it_sum / (N + 1), 'solv_error': rs_sum / (N + 1), 'alpha': alpha,
             'seconds_solve': t_solve, 'seconds_gram_stage': t_gram}
    return grads, stats
