"""Matrix-free log det K~ by stochastic Lanczos quadrature.

The reference computes ``log_det_K`` from a dense Cholesky factor (models/interpolated_llgp.py:262-276),
which only exists for small n; its roadmap lists the Lanczos estimator (README.md:88-93).  MINRES is a
Lanczos process, so the block solver hands back the tridiagonal T of every right-hand side it solves
(lmc_minres_lanczos, include/lmc_b200.h) and for a Rademacher probe z

    z^T log(K~) z  ~=  ||z||^2 e1^T log(T) e1 ,        log det K~ = E[z^T log(K~) z].

The probes are the ones a gradient evaluation solves against anyway (stochastic_deriv.py:33-38), so the
estimate costs no additional product.
"""
import numpy as np
import scipy.linalg

from .iterative import fused_of


def quadrature_terms(tridiag, beta1, iters):
    """Per right-hand side: beta1^2 e1^T log(T_k) e1, k = min(iterations, recorded steps).

    tridiag: [P, K, 2] of (alfa_j, beta_{j+1}); beta1: [P]; iters: [P]."""
    tridiag = np.asarray(tridiag)
    out = np.zeros(len(beta1))
    for c in range(len(beta1)):
        k = int(min(iters[c], tridiag.shape[1]))
        if k == 0:
            continue
        theta, vecs = scipy.linalg.eigh_tridiagonal(tridiag[c, :k, 0], tridiag[c, :k - 1, 1])
        out[c] = beta1[c] ** 2 * np.sum(vecs[0] ** 2 * np.log(theta))
    return out


def stochastic_logdet(K, n_probes=16, steps=None, tol=1e-4, maxiter=None, probes=None):
    """Estimate log det K for a fused operator tree K (gen_grid_kernel's result).

    :param n_probes: number of Rademacher probes drawn from the global numpy RNG, like the reference's
        gradient probes (stochastic_deriv.py:35); ignored when ``probes`` ([N, n], entries +-1) is given.
    :param steps: Lanczos steps kept per probe (default: as many as the solve runs, at most 2000).
    :returns: (estimate, standard error over the probes, per-probe terms)
    """
    import torch
    fused = fused_of(K)
    if fused is None:
        raise ValueError('stochastic_logdet needs the fused device operator of gen_grid_kernel')
    n = K.shape[0]
    if probes is None:
        probes = np.random.randint(0, 2, (n_probes, n)) * 2.0 - 1.0
    probes = np.ascontiguousarray(probes, dtype=np.float64)
    maxiter = n if maxiter is None else maxiter
    steps = min(maxiter, 2000) if steps is None else steps
    R = torch.as_tensor(probes, device='cuda')
    _, iters, _, istop, tri, beta1 = fused.minres_lanczos_device(R, steps, tol=tol, maxiter=maxiter)
    terms = quadrature_terms(tri, beta1, iters)
    est = float(np.mean(terms))
    err = float(np.std(terms, ddof=1) / np.sqrt(len(terms))) if len(terms) > 1 else float('nan')
    return est, err, terms
