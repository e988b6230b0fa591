"""One IRLS (iteratively re-weighted least squares) inner step around the hot path, with
everything resident in HBM.

The reference has no such function: a glum-style solver calls ``X.matvec(beta)`` for the linear
predictor, computes the working weights and the score rows on the host, and then calls
``X.sandwich(d)`` and ``X.transpose_matvec(v)`` separately (callers of matrix_base.py:15-77) —
two passes over X and an n-vector host<->device round trip per iteration.  Here

    eta = X @ beta                                   (device, stays there)
    d, v = weights_fn(eta)                           (caller-supplied element-wise torch op)
    H, g = X.sandwich_and_transpose_matvec(d, v)     (ONE pass over the dense block)

and only the p x p Hessian and the length-p score ever need to leave the GPU.
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch

from . import _dev

WeightsFn = Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]]


def irls_step(X, beta, weights_fn: WeightsFn, offset: Optional[torch.Tensor] = None,
              rows=None, cols=None, stored_order: bool = False):
    """``(H, g, eta)`` with ``eta = X @ beta (+ offset)``, ``(d, v) = weights_fn(eta)``,
    ``H = X.T diag(d) X`` and ``g = X.T v`` (restricted to ``rows`` / ``cols`` like the
    reference's ``sandwich`` / ``transpose_matvec``).

    ``beta``: host array or CUDA tensor of length p.  ``weights_fn`` maps the CUDA tensor
    ``eta`` (length n, the matrix dtype) to two CUDA tensors of the same length and dtype —
    e.g. for logistic regression ``mu = sigmoid(eta); return mu * (1 - mu), y - mu``.
    Everything is returned on the device.

    ``stored_order=True`` (a ``RowSortedMatrix``): ``eta``, ``offset`` and whatever per-row data
    ``weights_fn`` closes over (the response ...) are in the STORED row order
    (``X.to_stored_order(y)``, once) - XᵀDX and Xᵀv do not depend on the row order, so the
    iteration then runs without the permutation of ``eta`` back to the caller's order and of
    ``d`` and ``v`` into the stored one (three n-vector gathers / scatters per step).  ``rows``
    stays in the caller's numbering.  The returned ``eta`` is in the stored order too
    (``X.from_stored_order``)."""
    tdt = _dev.torch_dtype(X.dtype)
    b = beta if _dev.is_dev(beta) else _dev.to_dev(np.asarray(beta), tdt)
    if b.dtype != tdt:
        b = b.to(tdt)
    if stored_order and hasattr(X, "to_stored_order"):
        inner = X.mat
        eta = inner.matvec(b, cols)
        if offset is not None:
            eta = eta + offset
        d, v = weights_fn(eta)
        if d.dtype != tdt or v.dtype != tdt:
            d, v = d.to(tdt), v.to(tdt)
        H, g = inner.sandwich_and_transpose_matvec(d.contiguous(), v.contiguous(),
                                                    X._rows_in(rows), cols)
        return H, g, eta
    eta = X.matvec(b, cols)
    if offset is not None:
        eta = eta + offset
    d, v = weights_fn(eta)
    if d.dtype != tdt or v.dtype != tdt:
        d, v = d.to(tdt), v.to(tdt)
    H, g = X.sandwich_and_transpose_matvec(d.contiguous(), v.contiguous(), rows, cols)
    return H, g, eta
