"""runlmc.util.inline_pool.  On the device backend the right-hand sides of a gradient evaluation
are solved as ONE multi-RHS solve, so this pool only serves operator trees the fused path does not
recognise (one task per right-hand side, like the reference's, inline_pool.py:16-19)."""


def _one_after_the_other(f, argument_tuples):
    return [f(*args) for args in argument_tuples]


class InlinePool:
    """`starmap` over a multiprocessing.Pool when one is given, else in this process."""

    def __init__(self, pool):
        self._pool = pool

    def starmap(self, f, ls):
        run = _one_after_the_other if not self._pool else self._pool.starmap
        return run(f, ls)

    def __del__(self):
        if self._pool:
            self._pool.close()
