"""LayerNorm — host entry points for K4 (replaces reference vit/kernels/layernorm.py:90-142)."""
from typing import Optional

import torch

from . import _lib


def layernorm(A: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float,
              out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Normalise the last dim of a contiguous (B, N, D) tensor: w * (x - mean) / sqrt(var + eps) + b.

    Biased variance, eps inside the square root — the arithmetic of the reference's
    ``layernorm_kernel`` (layernorm.py:51-85).  ``out_dtype`` (bf16 from an fp32 input) is an
    extension used by mixed-precision callers; by default the output has the input dtype.
    """
    assert A.is_contiguous(), 'Matrix is not contiguous'
    assert A.is_cuda, 'Matrix is not on GPU'
    assert len(A.shape) == 3, "Only 3 dimensional matrix is supported as input"
    dim = A.shape[-1]
    assert weight.numel() == dim and bias.numel() == dim, \
        f"Weight/bias size should match the embedding dimension, provided: {weight.shape}, {bias.shape}, {dim}"
    assert weight.dtype == A.dtype and bias.dtype == A.dtype, \
        f"Weight/bias dtype should match the input dtype, provided: {weight.dtype}, {bias.dtype}, {A.dtype}"

    out = torch.empty_like(A) if out_dtype is None else torch.empty(A.shape, device=A.device, dtype=out_dtype)
    rows = A.shape[0] * A.shape[1]
    _lib.call("vt_layernorm", A.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), rows,
              dim, dim, dim, float(eps), _lib.dtype_code(A), _lib.dtype_code(out), _lib.stream_ptr(A))
    return out


class LayerNormTriton(torch.nn.Module):
    """Drop-in for the reference module of the same name (layernorm.py:129-142); the name is kept so
    existing state-dicts and call sites keep working — the kernel underneath is CUDA, not Triton."""

    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.ones(self.dim))
        self.bias = torch.nn.Parameter(torch.zeros(self.dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return layernorm(x, self.weight, self.bias, self.eps)


LayerNormB200 = LayerNormTriton
