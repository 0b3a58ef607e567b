"""runlmc.linalg.sum_matrix: `SumMatrix` lives in operators.py with the other composite operators."""
from .operators import SumMatrix  # noqa: F401
