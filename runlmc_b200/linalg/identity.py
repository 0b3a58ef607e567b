"""runlmc.linalg.identity: `Identity` lives in operators.py with the other composite operators."""
from .operators import Identity  # noqa: F401
