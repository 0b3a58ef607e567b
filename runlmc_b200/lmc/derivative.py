class Derivative:
    """dL/dtheta = (d_normal_quadratic - d_logdet_K) / 2 (reference lmc/derivative.py:5-12)."""

    def derivative(self, dKdt):
        return 0.5 * (self.d_normal_quadratic(dKdt) - self.d_logdet_K(dKdt))

    def d_normal_quadratic(self, dKdt):
        raise NotImplementedError

    def d_logdet_K(self, dKdt):
        raise NotImplementedError
