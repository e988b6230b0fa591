"""Host-side mirror of the reference's plugin boundary (``tabmat.ext.{dense,sparse,categorical,split}``).

Same function names and argument meaning as the reference's Cython modules, but every array
argument is a CUDA ``torch.Tensor`` and every function launches sm_100a kernels through the
C-ABI in ``include/tabmat_b200.h``.  ``rows`` / ``cols`` may be ``None`` (= all), in which case
no index array is materialised.
"""
