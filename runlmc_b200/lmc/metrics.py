"""runlmc.lmc.metrics: `Metrics` is defined next to the service that fills it, stochastic_deriv.py."""
from .stochastic_deriv import Metrics  # noqa: F401
