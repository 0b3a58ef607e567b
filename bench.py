#!/usr/bin/env python
"""Benchmark of the SKI-LMC hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload E] [--impl reference]

One step = one block product K~ V over the workload's right-hand sides
(y + the Hutchinson probes).  `value` is MVM*RHS per second with V resident in
HBM; `e2e` is the same metric through the host-buffer C-ABI path (pinned host
-> device copy of V and device -> host copy of K~V inside the timed region).
The probes are sharded over the ranks (strong scaling: total work fixed); no
collective is on the MVM path.  Also reported: per-kernel-family device time
with the roofline of the dominant family and of the whole product, one full
stochastic gradient evaluation (solves + all partials), and the CPU oracle
timed on the host cores.

--impl reference times the reference's own CPU algorithm (the numpy/scipy
oracle port, one process per core like the reference's Pool.starmap) on a
bounded sample of the same workload.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault('OMP_NUM_THREADS', '1')   # the reference's own setting (benchlib/bench.py:7)

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from runlmc_b200 import synthetic  # noqa: E402

METRIC = 'ski_lmc_mvm_rhs_per_s'
UNIT = 'MVM*RHS/s'
CPL = {'A': 4, 'B': 8, 'C': 5, 'D': 2, 'E': 1.5}     # grid cells per (shortest) lengthscale
# hyper-parameters of the headline gradient row: kernels resolved by the grid and a noise level on which the
# reference's stopping rule converges below tol (picked with tools/conv_probe.py)
# (tools/conv_probe.py, profiles/r02_conv_probe.txt).  At n = 1M scipy's rule -- relative to ||A|| ||x|| with
# ||b|| = 1000 -- only reaches an absolute residual of 1e-4 when kappa_2 is O(10^2): same kernels as the
# product benchmark, noise variances scaled up (a low signal-to-noise model).
GRAD_PARAMS = {'E': dict(cpl=1.5, eps=1.0, noise_scale=300.0), 'D': dict(cpl=2.0, eps=1.0, noise_scale=100.0)}


def workload_desc(name, prob):
    """Identical in both arms (the driver compares `config` between them)."""
    return {'workload': '%s: D=%d n=%d grid=%s Q=%d probes=%d (SURVEY.md sec.8 config %s)' % (
        name, prob.D, prob.n, 'x'.join(map(str, prob.grid_sizes)), prob.Q, prob.N, name),
        'rhs_per_step': prob.N + 1,
        'cache': 'inputs of a step (%.2f GB) are larger than the 126 MB L2' % ((prob.N + 1) * prob.n * 8 / 1e9)}


# --------------------------------------------------------------------------
# CPU side (oracle port of the reference path)
# --------------------------------------------------------------------------
_CPU_OP = None
_CPU_K = 0


def _cpu_mvm(v):
    return _CPU_OP.matvec(v)


def _cpu_solve_capped(b):
    """_CPU_K iterations of the reference's solver on one right-hand side (one pool task per column, like
    pool.starmap(Iterative.solve, ...), stochastic_deriv.py:51-52)."""
    from oracle import lmc_oracle as orc
    t0 = time.perf_counter()
    orc.minres(_CPU_OP.matvec, b, 1e-10, _CPU_K)
    return time.perf_counter() - t0


def oracle_operator(prob):
    from oracle import lmc_oracle as orc
    spec = orc.KernelSpec(['rbf'] * prob.Q, [[g] for g in prob.gammas], prob.coreg_vecs,
                          prob.coreg_diags, prob.noise)
    return orc.build_operator(spec, prob.Xs, prob.grids, rep='sum')


def cpu_reference_rate(prob, steps, warmup, cores=None, budget_s=25.0, solve_iters=0, keep=False):
    """MVM/s of the oracle (= the reference's numpy/scipy arithmetic) with one
    worker process per core, each step a block of `cores` columns.  solve_iters > 0 also times that
    many MINRES iterations per column through the same pool (the reference's parallel mode)."""
    global _CPU_OP, _CPU_K
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    _CPU_OP = oracle_operator(prob)
    t_build = time.perf_counter() - t0
    cols = [prob.probes[i % prob.N] for i in range(cores)]
    ctx = mp.get_context('fork')
    out = {}
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(_cpu_mvm, cols)
        done = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_mvm, cols)
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    if solve_iters:
        _CPU_K = solve_iters
        with ctx.Pool(cores) as pool:           # forked after _CPU_K is set
            t0 = time.perf_counter()
            pool.map(_cpu_solve_capped, cols)
            dts = time.perf_counter() - t0
        out['minres_iter_rhs_per_s'] = cores * solve_iters / dts
        out['minres_sample'] = 'Pool(%d).map of %d-iteration MINRES solves, one column per core, %.1f s' % (
            cores, solve_iters, dts)
    if not keep:
        _CPU_OP = None
    out.update({'value': cores * done / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                'sample': '%d steps x %d columns (one per core) of the %d-column block; operator build %.1f s not timed'
                          % (done, cores, prob.N + 1, t_build), 'ms_per_step': 1e3 * dt / done, 'steps': done})
    return out


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, sampled from a host thread every few ms through
    NVML (in-process; `nvidia-smi` subprocess as a fallback) so that even a ~100 ms timed region
    holds several samples.  `window` = wall-clock bounds of the device-timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period, self.rows, self.stop_flag = index, period, [], False
        self.window = [None, None]
        self.nvml = self.h = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(',')]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        flags = [bool(r & n.nvmlClocksEventReasonHwSlowdown), bool(r & n.nvmlClocksEventReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksEventReasonSwThermalSlowdown), bool(r & n.nvmlClocksEventReasonSwPowerCap)]
        return [time.time(), mhz, self.max_mhz] + flags

    def _sample_smi(self):
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                              '--format=csv,noheader,nounits'], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        if not out:
            return None
        f = [x.strip() for x in out.split(',')]
        return [time.time(), float(f[0]), float(f[1])] + [x == 'Active' for x in f[2:6]]

    def run(self):
        while not self.stop_flag:
            try:
                row = self._sample_nvml() if self.nvml else self._sample_smi()
                if row:
                    self.rows.append(row)
            except Exception:
                pass
            time.sleep(self.period if self.nvml else 0.2)

    def summary(self):
        self.stop_flag = True
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        inside = [r for r in self.rows if self.window[0] is not None and self.window[0] <= r[0] <= self.window[1]]
        use = inside or self.rows
        sm = sorted(r[1] for r in use)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[3 + i] for r in use)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_min_mhz': sm[0], 'sm_max_mhz': self.rows[0][2], 'reasons': reasons,
                'samples': len(self.rows), 'samples_in_timed_region': len(inside),
                'source': 'nvml' if self.nvml else 'nvidia-smi',
                'note': 'median/min/reasons over the samples taken inside the device-timed region '
                        '(every %.0f ms from a host thread); all samples if the region held none'
                        % (1e3 * (self.period if self.nvml else 0.2))}


# --------------------------------------------------------------------------
# roofline bookkeeping (DESIGN.md "Algorithmic bytes")
# --------------------------------------------------------------------------
def family_bytes(family, prob, P, bins, gpitch, rows=False):
    """Algorithmic HBM bytes of one step for a kernel family: the vectors and
    coordinates a stage must read/write plus its grid-side slabs once in and once out."""
    n, d, D = prob.n, prob.ndim, prob.D
    pairs = (P + 1) // 2
    slab = 16.0 * D * pairs                     # bytes per complex element over all (pair, output) slabs
    if family == 'to_grid':
        return 8.0 * n * P + 8.0 * d * n + slab * gpitch
    if family == 'from_grid':
        return 16.0 * n * P + 8.0 * d * n + slab * gpitch
    if family == 'other' and rows:
        # point-major blocks whose gather cannot write rows itself (no such geometry among the benchmark
        # workloads): one transposing pass, sorted column-major result + the caller's rows of X in, rows of Y out
        return 24.0 * n * P + 4.0 * n
    if family == 'other':
        # the two permutation passes of the caller-order product (caller -> sorted before the scatter,
        # sorted -> caller after the gather): each reads and writes the block once, plus the index
        return 2 * (16.0 * n * P + 4.0 * n)
    if d == 2:
        # fused 2-D path: row transforms G <-> S_T[ky][x] (only the m0 non-zero / kept x columns are
        # stored), column transform + mix + inverse in place on S_T
        st = float(bins) / (2 * prob.grid_sizes[0]) * (-(-prob.grid_sizes[0] // 8) * 8)
        if family in ('fft_fwd_contig', 'fft_inv_contig'):
            return slab * (gpitch + st)
        if family == 'mix':
            return slab * 2 * st
        return 0.0
    if family in ('fft_fwd_contig', 'fft_inv_contig'):
        return slab * (gpitch + bins)
    if family in ('fft_fwd_strided', 'fft_inv_strided'):
        return slab * 1.5 * bins
    if family == 'mix':
        return slab * 2 * (bins if bins > 8192 else gpitch)   # short 1-D lines are transformed in the grid slabs
    return 0.0


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(workload, family, pairs_per_launch):
    """Measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one launch of the family's
    kernel: profiles/ncu_traffic.json holds bytes per RHS pair from the committed `ncu --set full`
    capture of this workload; a launch moves that times the pairs it processes."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(p):
        per_pair = json.load(open(p)).get(workload, {}).get(family)
        if per_pair:
            return per_pair * pairs_per_launch
    return None


# --------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """One process per GPU: run on the CPUs next to this rank's GPU (NVML's ideal affinity) so that the pinned
    host buffers of the end-to-end leg are allocated on the GPU's own NUMA node instead of wherever torchrun
    started the rank.  Returns the number of CPUs bound to, or None if the platform does not allow it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = index
        if vis:
            ids = [v.strip() for v in vis.split(',')]
            if index < len(ids) and ids[index].isdigit():
                phys = int(ids[index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def rfft_bins(op):
    """Number of rfft bins m~_c of the embedding: m~_1 ... (m~_P / 2 + 1)."""
    mt = [1 << int(2 * m - 1).bit_length() for m in op.grid_sizes]
    out = mt[-1] // 2 + 1
    for v in mt[:-1]:
        out *= v
    return out


def parity_check(op, ref, Vh, OUT, tol=1e-10):
    """The timed operator's own output against the oracle on the first pair and the odd last column."""
    cols = sorted({0, min(1, len(Vh) - 1), len(Vh) - 1})
    errs = []
    for c in cols:
        want = ref.matvec(Vh[c])
        got = OUT[c].cpu().numpy()
        errs.append(float(np.linalg.norm(got - want) / np.linalg.norm(want)))
    return {'columns': cols, 'max_rel_err_vs_oracle': max(errs), 'tolerance': tol, 'ok': max(errs) <= tol,
            'oracle': 'oracle/lmc_oracle.py (numpy/scipy restatement pinned to the reference)'}


def timed_product(product, steps, warmup, barrier, min_warm_s=0.5):
    import torch
    t_w = time.time()
    while True:                                   # warm-up: >= W steps and long enough for the clocks to settle
        for _ in range(warmup):
            product()
        torch.cuda.synchronize()
        if time.time() - t_w > min_warm_s:
            break
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        product()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def gradient_row(op, prob, rank, world, dev, barrier, peak, label, note):
    """One full stochastic gradient evaluation (solves + all partials) with the operator's current
    hyper-parameters."""
    import torch
    import torch.distributed as dist
    from runlmc_b200.distributed import sharded_gradient
    # kappa_2 estimate: lambda_max by power iteration, lambda_min >= the smallest noise variance
    v = torch.as_tensor(prob.probes[:1].copy(), device=dev)
    lam = 0.0
    for _ in range(30):
        w = op.mvm_device(v)
        lam = float(torch.linalg.norm(w))
        v = (w / lam).contiguous()
    kappa2 = lam / float(prob.noise.min())
    barrier()
    t0 = time.perf_counter()
    grads, stats = sharded_gradient(op, prob.y, prob.probes, None, prob.coreg_vecs,
                                    prob.coreg_mats(), tol=1e-4, rank=rank, world=world)
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    flat = np.concatenate([np.ravel(g) for g in grads[0]] + [np.ravel(g) for g in grads[1]] +
                          [np.ravel(g) for g in grads[2]] + [np.ravel(grads[3])])
    it_rate = stats['iterations'] * (prob.N + 1) / dt
    return {'row': label, 'note': note, 'grad_evals_per_s': 1.0 / dt, 'seconds': dt,
            'mean_minres_iterations': stats['iterations'], 'mean_final_residual': stats['solv_error'],
            'tol': 1e-4, 'converged': bool(stats['solv_error'] < 1e-4),
            'seconds_solves_this_rank': stats['seconds_solve'], 'seconds_gram_stage_this_rank': stats['seconds_gram_stage'],
            'hyperparameters': int(flat.size), 'minres_iter_rhs_per_s': it_rate,
            'lengthscales_in_grid_cells': [float(1.0 / np.sqrt(g) * (max(prob.grid_sizes) - 1)) for g in prob.gammas],
            'noise': [float(prob.noise.min()), float(prob.noise.max())],
            'kappa2_upper_estimate': kappa2, 'lambda_max': lam,
            'gradient_l2': float(np.linalg.norm(flat)), 'gradient_checksum': float(flat.sum()),
            'gradient_values': [float(v) for v in flat],
            'includes': 'host->device copies of y/probes, N+1 MINRES solves (tol 1e-4, reference stopping rules), '
                        'all partial derivatives, allreduce',
            # MINRES against B_minres_iter = 120 n bytes per iteration and RHS
            'roofline_minres': {'bound': 'hbm', 'alg_bytes_per_iter_rhs': 120.0 * prob.n,
                                'achieved': 120.0 * prob.n * it_rate / world / 1e9, 'peak': peak, 'unit': 'GB/s',
                                'frac': 120.0 * prob.n * it_rate / world / 1e9 / peak, 'note': 'per GPU'}}


def small_config_row(name, peak):
    """A secondary workload on ONE GPU: block-product rate and MINRES iteration rate of the device path,
    with the CPU port's serial MINRES iteration rate beside it (configs A/B/C are launch bound, D is the
    HBM-bound one)."""
    import torch
    from oracle import lmc_oracle as orc
    from runlmc_b200.fused import FusedLMC
    from runlmc_b200.kern import RBF
    from runlmc_b200 import _native as nat
    prob = synthetic.make_problem(name, seed=1234, cells_per_lengthscale=CPL[name])
    op = FusedLMC(prob.Xs, prob.grids)
    op.set_kernels([RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs, prob.coreg_diags)
    Vh = np.ascontiguousarray(np.vstack([prob.y[None, :], prob.probes]))
    P = Vh.shape[0]
    V = torch.as_tensor(Vh, device='cuda')
    OUT = torch.empty_like(V)
    steps = 20 if prob.n >= 100000 else 200
    X = V.t().contiguous()                 # point-major block, like the headline row
    Y = torch.empty_like(X)
    ms = timed_product(lambda: op.matmat_device(X, Y), steps, 3, torch.cuda.synchronize, min_warm_s=0.1) / steps
    ms_cols = timed_product(lambda: op.mvm_device(V, OUT), steps, 3, torch.cuda.synchronize, min_warm_s=0.1) / steps
    perm = torch.as_tensor(op.perm().astype(np.int64), device='cuda')
    Vs = V[:, perm].contiguous()
    for _ in range(3):
        op.mvm_sorted_device(Vs, OUT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        op.mvm_sorted_device(Vs, OUT)
    e1.record()
    torch.cuda.synchronize()
    ms_sorted = e0.elapsed_time(e1) / steps
    bins = nat.lib.lmc_op_embed_bins(op._h)
    b_mvm = 16.0 * prob.n * P + 8.0 * prob.ndim * prob.n + 8.0 * prob.Q * bins + 8.0 * prob.D
    # MINRES: K iterations on the whole block (no column converges that early), launches counted
    K = 50
    op.minres_device(V, tol=1e-4, maxiter=5, check_every=10 ** 6)
    torch.cuda.synchronize()
    l0 = nat.lib.lmc_launch_count()
    t0 = time.perf_counter()
    _, iters, _, _ = op.minres_device(V, tol=1e-4, maxiter=K, check_every=10 ** 6)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = int(nat.lib.lmc_launch_count() - l0)
    done = float(np.sum(iters))
    # CPU port, serial, same iteration count on a few columns (bounded)
    ref = oracle_operator(prob)
    ncpu = 2 if prob.n >= 100000 else 4
    t0 = time.perf_counter()
    for c in range(ncpu):
        orc.minres(ref.matvec, Vh[c], 1e-10, K)
    dcpu = time.perf_counter() - t0
    par = parity_check(op, ref, Vh, op.matmat_device(X).t())
    return {'workload': workload_desc(name, prob)['workload'], 'mvm_rhs_per_s': P / ms * 1e3, 'ms_per_step': ms,
            'ms_per_step_column_major': ms_cols, 'ms_per_step_sorted_order': ms_sorted,
            'roofline_mvm': {'bound': 'hbm', 'alg_bytes_per_step': b_mvm, 'achieved': b_mvm / ms / 1e6, 'peak': peak,
                             'unit': 'GB/s', 'frac': b_mvm / ms / 1e6 / peak,
                             'frac_sorted_order': b_mvm / ms_sorted / 1e6 / peak},
            'minres_iter_rhs_per_s': done / dt, 'minres_iterations_timed': K,
            'minres_kernel_launches_per_iteration': launches / K,
            'minres_host_launches_per_iteration': 'one cudaGraphLaunch + one 4-byte stop-flag copy',
            'cpu_port_minres_iter_rhs_per_s': ncpu * K / dcpu, 'cpu_port_cores': 1,
            'cpu_port_sample': '%d columns x %d iterations, serial' % (ncpu, K), 'parity': par}


def run_own(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    prob = synthetic.make_problem(args.workload, seed=1234, cells_per_lengthscale=CPL[args.workload])
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # before CUDA is initialised (fork); also times capped MINRES solves through the pool
        cpu = cpu_reference_rate(prob, steps=3, warmup=1, solve_iters=8 if prob.n >= 100000 else 40, keep=True)
    ref = _CPU_OP if _CPU_OP is not None else oracle_operator(prob)
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None   # before any pinned host buffer is allocated
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from runlmc_b200 import _native as nat
    from runlmc_b200.fused import FusedLMC
    from runlmc_b200.distributed import shard_bounds
    dev = torch.device('cuda', local)
    from runlmc_b200.kern import RBF
    op = FusedLMC(prob.Xs, prob.grids)                       # point sort on the device
    op.set_kernels([RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs,
                   prob.coreg_diags)                         # kernel values evaluated on the device
    lo, hi = shard_bounds(prob.N, rank, world)
    rows = [prob.y[None, :]] if rank == 0 else []
    rows.append(prob.probes[lo:hi])
    Vh = np.ascontiguousarray(np.vstack(rows))
    P = Vh.shape[0]
    V = torch.as_tensor(Vh, device=dev)
    OUT = torch.empty_like(V)
    total_units = prob.N + 1
    # The timed block is the [n, P] argument of the reference's Matrix.matmat as numpy lays it out (C order,
    # point-major: lmc_mvm_rows); --layout cols times P contiguous columns instead (lmc_mvm).  Both are in the
    # caller's point order and both are reported.
    if args.layout == 'rows':
        X = V.t().contiguous()
        Y = torch.empty_like(X)
        OUT_cols = Y.t()                      # view: column c of the block

        def product():
            op.matmat_device(X, Y)

        def other():
            op.mvm_device(V, OUT)
    else:
        OUT_cols = OUT

        def product():
            op.mvm_device(V, OUT)
        other = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        product()
    torch.cuda.synchronize()
    l0 = nat.lib.lmc_launch_count()
    sampler.window[0] = time.time()
    ms = timed_product(product, args.steps, args.warmup, barrier)
    sampler.window[1] = time.time()
    launches = int(nat.lib.lmc_launch_count() - l0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_step = ms / args.steps
    value = total_units * args.steps / (ms / 1e3)
    parity = parity_check(op, ref, Vh, OUT_cols)
    pe = torch.tensor([parity['max_rel_err_vs_oracle']], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pe, op=dist.ReduceOp.MAX)
    parity['max_rel_err_vs_oracle'] = float(pe.item())
    parity['ok'] = parity['max_rel_err_vs_oracle'] <= parity['tolerance']
    parity['config'] = args.workload
    parity['note'] = 'columns of every rank\'s shard of the timed block, max over ranks'
    layouts = {'timed': 'point-major [n, P] (lmc_mvm_rows)' if args.layout == 'rows' else 'P columns (lmc_mvm)',
               'ms_per_step': ms_step}
    if other is not None:
        other_steps = max(3, args.steps // 2)
        t = torch.tensor([timed_product(other, other_steps, 3, barrier, min_warm_s=0.1)], dtype=torch.float64,
                         device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        layouts['column_major_ms_per_step'] = float(t.item()) / other_steps
        layouts['column_major_value'] = total_units * other_steps / (float(t.item()) / 1e3)
        layouts['max_abs_diff_between_layouts'] = float((OUT - OUT_cols).abs().max().item())

    # ---- end to end through host buffers: the public host API FusedLMC.mvm_into (C ABI
    # lmc_mvm_host) on pinned buffers; H2D copy of V and D2H copy of K~V are inside the timed
    # region of every step ----
    e2e_steps = max(1, min(args.steps, 5))
    Vp = torch.as_tensor(Vh).pin_memory()
    Op = torch.empty_like(Vp).pin_memory()
    Vpn, Opn = Vp.numpy(), Op.numpy()
    op.mvm_into(Vpn, Opn)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        op.mvm_into(Vpn, Opn)          # returns after the last D2H copy has landed
    torch.cuda.synchronize()
    dt_e2e = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    t = torch.tensor([dt_e2e * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = total_units * e2e_steps / (float(t.item()) / 1e3)
    e2e_check = float(np.abs(Opn - OUT_cols.cpu().numpy()).max())
    e2e_parity = parity_check(op, ref, Vh, torch.as_tensor(Opn))
    clocks = sampler.summary()
    cnt = torch.tensor([float(P)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    io_bytes = int(cnt.item()) * prob.n * 8

    # ---- per-family device time (separate pass so events do not perturb `value`) ----
    l1 = nat.lib.lmc_launch_count()
    nat.profile_begin()
    for _ in range(args.steps):
        product()
    prof = nat.profile_end()
    launches = int(nat.lib.lmc_launch_count() - l1)          # kernels of exactly `steps` products
    bins, gpitch = nat.lib.lmc_op_embed_bins(op._h), nat.lib.lmc_op_grid_cells(op._h)
    peak, peak_src = peaks()
    tot = sum(v[0] for v in prof.values()) or 1.0
    fams = []
    for k, (m_, c_) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        b = family_bytes(k, prob, P, bins, gpitch, rows=args.layout == 'rows')
        fams.append({'family': k, 'ms_per_step': m_ / args.steps, 'launches_per_step': c_ // args.steps,
                     'share': m_ / tot, 'alg_gb_per_step': b / 1e9,
                     'alg_gbs': b / (m_ / args.steps) / 1e6 if m_ else None})
    top = fams[0]
    lpl = max(top['launches_per_step'], 1)
    roofline = {'kernel': top['family'], 'bound': 'hbm', 'achieved': top['alg_gbs'], 'peak': peak, 'unit': 'GB/s',
                'frac': (top['alg_gbs'] or 0.0) / peak,
                'traffic': ncu_traffic(args.workload, top['family'], ((P + 1) // 2) / lpl),
                'share_of_step': top['share'], 'alg_bytes_per_launch': top['alg_gb_per_step'] * 1e9 / lpl,
                'launch_ms': top['ms_per_step'] / lpl, 'peak_source': peak_src}
    b_mvm = 16.0 * prob.n * P + 8.0 * prob.ndim * prob.n + 8.0 * prob.Q * bins + 8.0 * prob.D
    roofline_mvm = {'bound': 'hbm', 'alg_bytes_per_step': b_mvm, 'achieved': b_mvm / ms_step / 1e6, 'peak': peak,
                    'unit': 'GB/s', 'frac': b_mvm / ms_step / 1e6 / peak,
                    'note': 'whole product vs B_mvm = 16nP + 8dn + 8Q*bins + 8D (SURVEY.md sec.8d), this rank'}

    # ---- the spectral stage against the fp64 roofline: per MVM
    # and RHS, pruned real FFTs 2 * D * 2.5 * M~ log2 M~ * 3/4 plus the mix 4 D^2 m~_c with m~_c the number
    # of rfft bins (a complex RHS pair shares the full set of bins) ----
    fp64 = np.zeros(1)
    nat.check(nat.lib.lmc_fp64_peak(nat.host_ptr(fp64)))
    spectral_ms = sum(f['ms_per_step'] for f in fams if f['family'].startswith('fft') or f['family'] == 'mix')
    flops_rhs = 2.0 * prob.D * 2.5 * bins * np.log2(bins) * 0.75 + 4.0 * prob.D ** 2 * rfft_bins(op)
    roofline_fp64 = {'bound': 'fp64', 'stage': 'fft + mix + inverse fft', 'alg_flops_per_step': flops_rhs * P,
                     'achieved': flops_rhs * P / (spectral_ms * 1e-3) / 1e12 if spectral_ms else None,
                     'peak': float(fp64[0]), 'unit': 'TFLOP/s',
                     'frac': flops_rhs * P / (spectral_ms * 1e-3) / 1e12 / float(fp64[0]) if spectral_ms else None,
                     'peak_source': 'measured (lmc_fp64_peak: DFMA microbenchmark on this GPU)'}

    # ---- one full stochastic gradient evaluation (solves + all partials): the headline row uses
    # hyper-parameters on which the reference's own stopping rule converges, the second
    # row keeps the ill-conditioned operator the block product is timed on ----
    grad = grad_ill = None
    if not args.no_grad:
        # untimed: the solver's grow-only workspace (10 P n doubles) and its pinned flags are allocated once per
        # model, like the operator's grid workspace during the product's warm-up
        Xw, _, _, _ = op.minres_device(V, tol=1e-4, maxiter=20)
        if P > 2:
            op.grad_grams_device(Xw[0], V[1:3].contiguous(), Xw[1:3].contiguous(), None)   # loads the Gram-stage kernels
        del Xw
        torch.cuda.synchronize()
        gp = GRAD_PARAMS.get(args.workload)
        if gp:
            pc = synthetic.make_problem(args.workload, seed=1234, cells_per_lengthscale=gp['cpl'], eps=gp['eps'],
                                        noise_scale=gp['noise_scale'])
            op.set_kernels([RBF(g) for g in pc.gammas], pc.coreg_mats(), pc.noise, pc.coreg_vecs, pc.coreg_diags)
            grad = gradient_row(op, pc, rank, world, dev, barrier, peak, 'converging',
                                'lengthscales of %g..%g grid cells, noise ~ %g / Gamma(1 + 1/eps, 1) with eps = %g '
                                '(eps: the reference benchmark\'s noise parameter, benchlib/bench.py:111-115)'
                                % (gp['cpl'], 4 * gp['cpl'], gp['noise_scale'], gp['eps']))
            op.set_kernels([RBF(g) for g in prob.gammas], prob.coreg_mats(), prob.noise, prob.coreg_vecs,
                           prob.coreg_diags)
        if not args.no_ill:
            grad_ill = gradient_row(op, prob, rank, world, dev, barrier, peak, 'ill_conditioned',
                                    'the operator the block product is timed on (eps = 0.1): scipy\'s own stopping '
                                    'rule ends above tol, the reference would log the solve as failed')
        if grad is None:
            grad, grad_ill = grad_ill, None
        if cpu is not None and grad is not None:
            # the reference's path for the same evaluation: N+1 solves x mean iterations on all host cores at
            # the MEASURED rate of capped Pool solves (partials not counted)
            rate = cpu.get('minres_iter_rhs_per_s')
            if rate:
                cpu_s = (prob.N + 1) * grad['mean_minres_iterations'] / rate
                cpu['grad_eval_seconds_extrapolated'] = cpu_s
                cpu['grad_evals_per_s_extrapolated'] = 1.0 / cpu_s
                cpu['grad_extrapolation'] = '(N+1) x mean iterations of the headline gradient row / measured CPU ' \
                                            'MINRES iteration rate'
    extra = []
    if rank == 0 and world == 1 and not args.no_configs:
        for name in ('D', 'B', 'C', 'A'):
            if name != args.workload:
                extra.append(small_config_row(name, peak))
    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_desc(args.workload, prob),
                'timing': 'CUDA events on the launching stream, max over ranks',
                'e2e': {'value': e2e_val, 'unit': UNIT, 'h2d_bytes_per_step': io_bytes, 'd2h_bytes_per_step': io_bytes,
                        'steps': e2e_steps, 'api': 'FusedLMC.mvm_into -> lmc_mvm_host (pinned host buffers, '
                        'chunked copy/compute/copy pipeline)', 'cpus_bound_next_to_gpu': numa,
                        'max_abs_diff_vs_resident': e2e_check,
                        'max_rel_err_vs_oracle': e2e_parity['max_rel_err_vs_oracle']},
                'gpu_launches': launches, 'clocks': clocks, 'parity': parity, 'layouts': layouts,
                'roofline': roofline,
                'roofline_mvm': roofline_mvm, 'roofline_fp64': roofline_fp64, 'kernel_families': fams,
                'gradient': grad, 'gradient_ill_conditioned': grad_ill, 'configs': extra}
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    prob = synthetic.make_problem(args.workload, seed=1234, cells_per_lengthscale=CPL[args.workload])
    r = cpu_reference_rate(prob, steps=args.steps, warmup=args.warmup, budget_s=120.0)
    line = {'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT,
            'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': r['steps'], 'warmup': args.warmup,
            'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': workload_desc(args.workload, prob),
            'timing': 'host wall clock around Pool.map of the CPU products',
            'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--workload', default='E', choices=sorted(CPL))
    ap.add_argument('--layout', default='rows', choices=['rows', 'cols'],
                    help='timed block: point-major [n, P] (numpy C order of the matmat argument) or P columns')
    ap.add_argument('--no-grad', action='store_true', help='skip the gradient-evaluation leg')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline leg')
    ap.add_argument('--no-ill', action='store_true', help='skip the ill-conditioned gradient row')
    ap.add_argument('--no-configs', action='store_true', help='skip the secondary workload rows')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'own' else args.warmup
    # stdout carries exactly ONE line (the JSON): anything libraries print there while we run
    # (e.g. NCCL's version banner) is diverted to stderr at the file-descriptor level
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out):
            if args.impl == 'reference':
                run_reference(args)
            else:
                run_own(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith('{')]
    if lines:
        print(lines[-1], flush=True)


if __name__ == '__main__':
    main()
