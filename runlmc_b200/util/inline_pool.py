"""runlmc.util.inline_pool.  On the device backend the right-hand sides of a gradient evaluation
are solved as ONE multi-RHS solve: `starmap(Iterative.solve, tasks)` over one shared operator is
recognised and batched; anything else runs one task after the other (or on the multiprocessing pool)
like the reference's (inline_pool.py:16-19)."""


def _one_after_the_other(f, argument_tuples):
    return [f(*args) for args in argument_tuples]


def _batched_solves(f, ls):
    """The reference's batching seam (stochastic_deriv.py:39-52, interpolated_llgp.py:394):
    `starmap(Iterative.solve, [(K, rhs, verbose, minres, tol), ...])` with ONE shared operator is one
    multi-right-hand-side device solve.  Returns the list aligned with `ls` (x, or (x, iterations,
    residual) for the verbose tasks), or None when the tasks are not of that form."""
    from ..approx.iterative import Iterative, solve_block
    if f is not Iterative.solve or len(ls) < 2:
        return None
    norm = []
    for args in ls:
        args = tuple(args)
        if not 2 <= len(args) <= 5:
            return None
        K, y = args[0], args[1]
        verbose, minres, tol = (args[2:] + (False, True, 1e-4)[len(args) - 2:])
        norm.append((K, y, bool(verbose), bool(minres), float(tol)))
    K0, _, _, minres0, tol0 = norm[0]
    if any(a[0] is not K0 or a[3] != minres0 or a[4] != tol0 for a in norm):
        return None
    if not hasattr(K0, '_apply_dev'):
        return None
    import numpy as np
    RHS = np.array([np.asarray(a[1], dtype=np.float64).reshape(-1) for a in norm])
    X, iters, resid, istop = solve_block(K0, RHS, tol=tol0, minres=minres0,
                                         preconditioner=getattr(K0, 'preconditioner', None))
    Iterative.report(K0, resid, istop, tol0, minres0)
    return [(x, int(it), float(r)) if a[2] else x for a, x, it, r in zip(norm, X, iters, resid)]


class InlinePool:
    """`starmap` over a multiprocessing.Pool when one is given, else in this process."""

    def __init__(self, pool):
        self._pool = pool

    def starmap(self, f, ls):
        ls = list(ls)
        batched = _batched_solves(f, ls)
        if batched is not None:
            return batched
        run = _one_after_the_other if not self._pool else self._pool.starmap
        return run(f, ls)

    def __del__(self):
        if self._pool:
            self._pool.close()
