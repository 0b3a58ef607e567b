"""Mirror of runlmc/util/inline_pool.py.  On the device backend the right-hand
sides of a gradient evaluation are solved as ONE multi-RHS MINRES, so the pool
is only used for operator trees the fused path does not recognise."""


class InlinePool:
    """:param pool: a multiprocessing.Pool or None (serial)."""

    def __init__(self, pool):
        self._pool = pool

    def starmap(self, f, ls):
        if self._pool:
            return self._pool.starmap(f, ls)
        return [f(*x) for x in ls]

    def __del__(self):
        if self._pool:
            self._pool.close()
