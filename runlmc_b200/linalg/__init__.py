from .matrix import Matrix
from .toeplitz import Toeplitz
from .bttb import BTTB
from .kronecker import Kronecker
from .sum_matrix import SumMatrix
from .diag import Diag
from .numpy_matrix import NumpyMatrix
from .identity import Identity
from .composition import Composition
from .block_diag import BlockDiag
from .block_matrix import SymmSquareBlockMatrix

__all__ = ['Matrix', 'Toeplitz', 'BTTB', 'Kronecker', 'SumMatrix', 'Diag', 'NumpyMatrix',
           'Identity', 'Composition', 'BlockDiag', 'SymmSquareBlockMatrix']
