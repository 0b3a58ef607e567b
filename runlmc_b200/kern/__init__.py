"""Stationary kernels k(r) on distances and their parameter gradients.

Plain-numpy restatement of the formulas of runlmc/kern/{rbf,matern32,
std_periodic}.py without the paramz parameter plumbing (paramz is not a
dependency here).  They only produce the O(m) kernel values on the inducing
grid that are uploaded to the device operator each optimiser step."""
import numpy as np


class StationaryKern:
    def __init__(self, name, active_dims=None):
        self.name = name
        self.active_dims = active_dims

    def from_dist(self, dists):
        raise NotImplementedError

    def kernel_gradient(self, dists):
        raise NotImplementedError

    def param_values(self):
        raise NotImplementedError

    def update_gradient(self, grad):
        self.gradient = list(grad)


class RBF(StationaryKern):
    """k(r) = exp(-gamma r^2 / 2)   (kern/rbf.py:39-54)"""

    def __init__(self, inv_lengthscale=1, name='rbf', active_dims=None):
        super().__init__(name, active_dims)
        self.inv_lengthscale = float(inv_lengthscale)

    def from_dist(self, dists):
        return np.exp(-0.5 * np.square(dists) * self.inv_lengthscale)

    def kernel_gradient(self, dists):
        sq = np.square(dists)
        return [np.exp(-0.5 * sq * self.inv_lengthscale) * -0.5 * sq]

    def param_values(self):
        return [self.inv_lengthscale]


class Matern32(StationaryKern):
    """k(r) = (1 + sqrt(3) gamma r) exp(-sqrt(3) gamma r)   (kern/matern32.py:39-57)"""

    def __init__(self, inv_lengthscale=1, name='matern32', active_dims=None):
        super().__init__(name, active_dims)
        self.inv_lengthscale = float(inv_lengthscale)

    def from_dist(self, dists):
        s = dists * np.sqrt(3) * self.inv_lengthscale
        return (1 + s) * np.exp(-s)

    def kernel_gradient(self, dists):
        s = dists * np.sqrt(3) * self.inv_lengthscale
        ds = dists * np.sqrt(3)
        e = np.exp(-s)
        return [(1 + s) * (e * -ds) + ds * e]

    def param_values(self):
        return [self.inv_lengthscale]


class StdPeriodic(StationaryKern):
    """k(r) = exp(-gamma/2 sin^2(pi r / T))   (kern/std_periodic.py:44-67)"""

    def __init__(self, inv_lengthscale=1, period=1, name='std_periodic', active_dims=None):
        super().__init__(name, active_dims)
        self.inv_lengthscale = float(inv_lengthscale)
        self.period = float(period)

    def from_dist(self, dists):
        sn = np.sin((np.pi / self.period) * dists)
        return np.exp(-0.5 * np.square(sn) * self.inv_lengthscale)

    def kernel_gradient(self, dists):
        sc = np.pi / self.period * dists
        sn = np.sin(sc)
        dsn = np.cos(sc) * sc
        dsn = dsn * (-1 / self.period * self.inv_lengthscale)
        sq = np.square(sn)
        e = np.exp(-0.5 * sq * self.inv_lengthscale)
        return [e * -0.5 * sq, e * -1 * sn * dsn]

    def param_values(self):
        return [self.inv_lengthscale, self.period]
