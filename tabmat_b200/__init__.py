"""tabmat_b200 — B200-native (sm_100a) sandwich / matvec / transpose_matvec path behind
tabmat's ``MatrixBase`` API.

Drop-in for the hot path of Quantco/tabmat: ``DenseMatrix``, ``SparseMatrix``,
``CategoricalMatrix``, ``SplitMatrix``, ``StandardizedMatrix`` keep the reference's method
signatures and return conventions; the data lives in HBM and every numeric method launches
hand-written CUDA kernels through the C-ABI in ``include/tabmat_b200.h``.
There is no CPU fallback.
"""

from ._lib import (  # noqa: F401
    TabmatB200Error,
    deterministic,
    launch_count,
    reset_launch_count,
    set_deterministic,
)
from .categorical_matrix import CategoricalMatrix
from .constructor import from_csc, from_df, from_pandas
from .dense_matrix import DenseMatrix
from .irls import irls_step
from .matrix_base import MatrixBase
from .row_order import RowSortedMatrix
from .sparse_matrix import SparseMatrix
from .split_matrix import SplitMatrix, as_tabmat, hstack
from .standardized_mat import StandardizedMatrix

__version__ = "0.1.0"

__all__ = [
    "DenseMatrix",
    "MatrixBase",
    "StandardizedMatrix",
    "SparseMatrix",
    "SplitMatrix",
    "RowSortedMatrix",
    "CategoricalMatrix",
    "as_tabmat",
    "hstack",
    "from_csc",
    "from_df",
    "from_pandas",
    "irls_step",
    "set_deterministic",
]
