"""runlmc.linalg.block_diag: `BlockDiag` lives in operators.py with the other composite operators."""
from .operators import BlockDiag  # noqa: F401
