class Metrics:
    """Per-iteration statistics (reference runlmc/lmc/metrics.py:4-10)."""

    def __init__(self):
        self.iterations = []
        self.grad_norms = []
        self.grad_error = []
        self.solv_error = []
        self.log_likely = []
