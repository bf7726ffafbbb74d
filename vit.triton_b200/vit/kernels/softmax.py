"""Row softmax — host entry point for K6 (replaces reference vit/kernels/softmax.py:36-74)."""
import torch

from . import _lib


def softmax(A: torch.Tensor) -> torch.Tensor:
    """Softmax over the last axis of a (B, N, D) tensor (max-subtracted, like softmax.py:26-31)."""
    assert A.is_cuda, "Input is not on GPU"
    assert len(A.shape) == 3, f"Input needs to be 3 dimensional, provided: {A.shape}"
    if not A.is_contiguous():
        A = A.contiguous()
    batch, rows, cols = A.shape
    out = torch.empty_like(A)
    _lib.call("vt_softmax", A.data_ptr(), out.data_ptr(), batch * rows, cols, cols, _lib.dtype_code(A),
              _lib.stream_ptr(A))
    return out
