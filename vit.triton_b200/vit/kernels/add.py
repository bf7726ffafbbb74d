"""Element-wise addition — host entry point for K5 (replaces reference vit/kernels/add.py:67-104)."""
import torch

from . import _lib


def add(input1: torch.Tensor, input2: torch.Tensor) -> torch.Tensor:
    """out = input1 + input2 for two contiguous (B, N, D) tensors of the same shape.

    Same contract and error wording as the reference's ``add_triton`` (add.py:82-84); the result is
    a fresh tensor, the inputs are untouched.
    """
    assert input1.is_cuda and input2.is_cuda, "Input matrix needs to be on GPU"
    assert input1.is_contiguous() and input2.is_contiguous(), "Input matrix needs to be contiguous"
    assert len(input1.shape) == 3, f"Only 3 dimensional input shapes are supported, provided: {input1.shape}"
    assert input1.shape == input2.shape, f"Input shapes need to be same, provided {input1.shape}, {input2.shape}"
    assert input1.dtype == input2.dtype, f"Input dtypes need to be same, provided {input1.dtype}, {input2.dtype}"

    out = torch.empty_like(input1)
    _lib.call("vt_add", input1.data_ptr(), input2.data_ptr(), out.data_ptr(), input1.numel(),
              _lib.dtype_code(input1), _lib.stream_ptr(input1))
    return out
