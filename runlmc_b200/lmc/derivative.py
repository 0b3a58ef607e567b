"""runlmc.lmc.derivative: `Derivative` is defined next to its estimator in stochastic_deriv.py."""
from .stochastic_deriv import Derivative  # noqa: F401
