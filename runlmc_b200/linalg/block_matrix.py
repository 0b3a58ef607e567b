"""runlmc.linalg.block_matrix: `SymmSquareBlockMatrix` lives in operators.py with the other composite operators."""
from .operators import SymmSquareBlockMatrix  # noqa: F401
