"""runlmc.linalg.composition: `Composition` lives in operators.py with the other composite operators."""
from .operators import Composition  # noqa: F401
