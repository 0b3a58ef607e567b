"""runlmc.linalg.diag: `Diag` lives in operators.py with the other composite operators."""
from .operators import Diag  # noqa: F401
