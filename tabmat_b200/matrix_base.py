"""``MatrixBase``: the abstract API shared by every matrix class (reference: matrix_base.py:7-258).

Inputs may be numpy arrays (results come back as numpy arrays) or CUDA torch tensors (results
stay on the device)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Optional, Union

import numpy as np
import torch

from . import _dev


class MatrixBase(ABC):
    """Base class for all matrix classes. ``MatrixBase`` cannot be instantiated."""

    ndim = 2
    shape: tuple
    dtype: np.dtype

    @abstractmethod
    def matvec(self, other, cols=None, out=None):
        """self[:, cols] @ other[cols]; adds into ``out`` in place when given."""

    @abstractmethod
    def transpose_matvec(self, vec, rows=None, cols=None, out=None):
        """self[rows, cols].T @ vec[rows]; with ``out``: out[cols] += result."""

    @abstractmethod
    def sandwich(self, d, rows=None, cols=None):
        """(self[rows, cols].T * d[rows]) @ self[rows, cols]."""

    def sandwich_and_transpose_matvec(self, d, v, rows=None, cols=None):
        """``(X[rows, cols].T @ diag(d[rows]) @ X[rows, cols],  X[rows, cols].T @ v[rows])`` — the
        Hessian and the score of one IRLS step.  The reference's callers make the two calls
        separately (matrix_base.py:15-77: ``sandwich`` and ``transpose_matvec``); classes that
        can compute both in ONE pass over the matrix override this (``SplitMatrix``)."""
        return self.sandwich(d, rows, cols), self.transpose_matvec(v, rows, cols)

    def __matmul__(self, other):
        return self.matvec(other)

    @abstractmethod
    def getcol(self, i: int):
        pass

    @property
    def A(self) -> np.ndarray:
        return self.toarray()

    @abstractmethod
    def toarray(self) -> np.ndarray:
        pass

    def __rmatmul__(self, other):
        # other @ X = (X.T @ other.T).T   (matrix_base.py:97-112)
        if not hasattr(other, "T"):
            other = np.asarray(other)
        return self.transpose_matvec(other.T).T

    @abstractmethod
    def astype(self, dtype, order="K", casting="unsafe", copy=True):
        pass

    def _get_col_means(self, weights):
        """Weighted column means = transpose_matvec(weights) (matrix_base.py:118-120)."""
        return self.transpose_matvec(weights)

    @abstractmethod
    def _get_col_stds(self, weights, col_means):
        pass

    def standardize(self, weights, center_predictors: bool, scale_predictors: bool):
        """Return (StandardizedMatrix, column means, column stds) (matrix_base.py:128-167).

        col_means is zeros when not centering; col_stds is None when not scaling."""
        from .standardized_mat import StandardizedMatrix

        weights = np.asarray(weights) if not _dev.is_dev(weights) else _dev.to_host(weights)
        col_means = self._get_col_means(weights)
        if scale_predictors:
            col_stds = self._get_col_stds(weights, col_means)
            mult = _one_over_var_inf_to_val(col_stds, 1.0)
            if center_predictors:
                shifter = -col_means * mult
                out_means = col_means
            else:
                shifter = np.zeros_like(col_means)
                out_means = shifter
        else:
            col_stds = None
            if center_predictors:
                shifter = -col_means
                out_means = col_means
            else:
                shifter = np.zeros_like(col_means)
                out_means = shifter
            mult = None
        return StandardizedMatrix(self, shifter, mult), out_means, col_stds

    @abstractmethod
    def __getitem__(self, item):
        pass

    @abstractmethod
    def get_names(self, type: str = "column", missing_prefix: Optional[str] = None,
                  indices: Optional[list] = None) -> list:
        pass

    def set_names(self, names, type: str = "column"):
        pass

    @property
    def column_names(self):
        return self.get_names(type="column")

    @column_names.setter
    def column_names(self, names):
        self.set_names(names, type="column")

    @property
    def term_names(self):
        return self.get_names(type="term")

    @term_names.setter
    def term_names(self, names):
        self.set_names(names, type="term")

    # numpy must defer to us for ``v @ X`` (matrix_base.py:245)
    __array_priority__ = 11
    __array_ufunc__ = None


def _one_over_var_inf_to_val(arr: np.ndarray, val: float) -> np.ndarray:
    """1/arr, with ``val`` where |arr| < 1e-7 (matrix_base.py:248-258)."""
    zeros = np.where(np.abs(arr) < 1e-7)
    with np.errstate(divide="ignore"):
        one_over = 1 / arr
    one_over[zeros] = val
    return one_over


def _names_with_default(names, missing_prefix, indices):
    names = np.array(names, dtype=object)
    if indices is None:
        indices = list(range(len(names)))
    if missing_prefix is not None:
        default_names = np.array([f"{missing_prefix}{i}" for i in indices], dtype=object)
        mask = np.array([n is None for n in names], dtype=bool)
        names[mask] = default_names[mask]
    return names.tolist()


def _vec_in(vec, dtype: Optional[torch.dtype] = None):
    """(device tensor, came_from_host) for a vector / matrix argument."""
    if _dev.is_dev(vec):
        t = vec if dtype is None or vec.dtype == dtype else vec.to(dtype)
        return t.contiguous(), False
    return _dev.to_dev(vec, dtype), True


__all__ = ["MatrixBase", "_one_over_var_inf_to_val", "_names_with_default", "_vec_in", "Any", "Union"]
