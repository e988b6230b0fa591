"""Device plumbing: torch tensors own the HBM buffers, streams come from torch.

Public matrix methods accept either host data (numpy arrays / lists) or CUDA torch tensors:
host in -> host (numpy) out, device in -> device (torch) out.  Everything in between runs on
the device through the C-ABI kernels.
"""

from __future__ import annotations

from typing import Any, Optional

import numpy as np
import torch

_NP2T = {
    np.dtype(np.float32): torch.float32,
    np.dtype(np.float64): torch.float64,
    np.dtype(np.int32): torch.int32,
    np.dtype(np.int64): torch.int64,
}
_T2NP = {v: k for k, v in _NP2T.items()}
_SUF = {torch.float32: "f32", torch.float64: "f64"}


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "tabmat_b200 needs a CUDA device (sm_100a); there is no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def is_dev(x: Any) -> bool:
    return isinstance(x, torch.Tensor) and x.is_cuda


def np_dtype(t: torch.dtype) -> np.dtype:
    return _T2NP[t]


def torch_dtype(d) -> torch.dtype:
    if isinstance(d, torch.dtype):
        return d
    return _NP2T[np.dtype(d)]


def suffix(t: torch.dtype) -> str:
    try:
        return _SUF[t]
    except KeyError:
        raise TypeError(f"tabmat_b200 kernels support float32/float64, got {t}") from None


def to_dev(x: Any, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Host array / list / tensor -> contiguous CUDA tensor (no copy if already there)."""
    dev = require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if a.dtype == object:
            raise TypeError("object arrays are not supported")
        if not a.flags.writeable:
            a = a.copy()  # torch.from_numpy refuses read-only buffers
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.to(dev, non_blocking=False)
    return t.contiguous()


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def ret(t: torch.Tensor, as_host: bool):
    return to_host(t) if as_host else t


def idx32(x: Any) -> Optional[torch.Tensor]:
    """rows / cols restriction -> int32 CUDA tensor (None stays None)."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        t = x.to(torch.int32)
        return t.to(require_cuda()).contiguous() if not t.is_cuda else t.contiguous()
    a = np.asarray(x)
    if a.dtype == bool:
        a = np.flatnonzero(a)
    return to_dev(a.astype(np.int32, copy=False).reshape(-1))


_EMPTY_ANCHOR: dict = {}


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device address for the C-ABI.  ``None`` -> NULL, which the entry points read as "all
    rows / cols"; an EMPTY tensor (whose data_ptr is 0 too) gets the address of a one-element
    anchor instead, so that an empty restriction stays an empty restriction."""
    if t is None:
        return None
    if t.numel() == 0:
        key = (t.device.index, t.dtype)
        if key not in _EMPTY_ANCHOR:
            _EMPTY_ANCHOR[key] = torch.zeros(4, dtype=t.dtype, device=t.device)
        return _EMPTY_ANCHOR[key].data_ptr()
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def length(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else int(t.numel())
