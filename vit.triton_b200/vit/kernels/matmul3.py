"""Batched matmul — host entry point replacing reference vit/kernels/matmul3.py:111-156."""
import torch

from . import _lib


def matmul3(A: torch.Tensor, B: torch.Tensor, apply_scaling: bool = False,
            scale_factor: float = 1.0) -> torch.Tensor:
    """O[b] = s * A[b] @ B[b] for A (batch, seq_len, dim), B (batch, dim, dim_out); s = scale_factor
    when apply_scaling else 1.  Same argument checks as the reference (matmul3.py:123-128).

    The model's attention does not come through here (K3 fuses QK^T, softmax and PV); this entry
    point serves callers of the reference API and the exact-fp32 path.
    """
    assert len(A.shape) == 3, "First input matrix needs to have 3 dimensions (B, T, C)"
    assert len(A.shape) == len(B.shape), "Both matrix should be 3 dimensional"
    assert A.shape[2] == B.shape[1], f"Dimensions are not compatible for matrix multiplication, provided: {A.shape}, {B.shape}"
    assert A.shape[0] == B.shape[0], f"Batch sizes are not the same, provided: {A.shape}, {B.shape}"
    assert A.device == B.device and A.is_cuda, "Both matrix should be on GPU"
    assert A.is_contiguous(), "First matrix is not contiguous"
    assert B.is_contiguous(), "Second matrix is not contiguous"
    assert A.dtype == B.dtype, f"Input dtypes need to be same, provided {A.dtype}, {B.dtype}"

    batch, M, K = A.shape
    N = B.shape[-1]
    O = torch.empty((batch, M, N), device=A.device, dtype=A.dtype)
    if O.numel() == 0:
        return O
    # grid.z carries the batch: split very large batches
    step = 32768
    for z0 in range(0, batch, step):
        zb = min(step, batch - z0)
        _lib.call("vt_gemm_strided", A[z0:].data_ptr(), B[z0:].data_ptr(), O[z0:].data_ptr(), None,
                  M, N, K, zb, 1,
                  _lib.i64x4(M * K, 0, K, 1), _lib.i64x4(K * N, 0, N, 1), _lib.i64x4(M * N, 0, N, 1),
                  float(scale_factor) if apply_scaling else 1.0, 0, _lib.dtype_code(A),
                  _lib.stream_ptr(A))
    return O
