"""Batched matmul — host entry point replacing reference vit/kernels/matmul3.py:111-156."""
import os

import torch

from . import _lib, bgemm as bg


def matmul3(A: torch.Tensor, B: torch.Tensor, apply_scaling: bool = False,
            scale_factor: float = 1.0) -> torch.Tensor:
    """O[b] = s * A[b] @ B[b] for A (batch, seq_len, dim), B (batch, dim, dim_out); s = scale_factor
    when apply_scaling else 1.  Same argument checks as the reference (matmul3.py:123-128).

    Runs on the tensor cores (``vt_bgemm``, csrc/bgemm_sm100.cu) like the reference's ``tl.dot`` kernel
    (matmul3.py:97): bf16 operands whose rows are multiples of 8 elements are consumed IN PLACE — B as
    the (dim, dim_out) row-major matrix the caller passes, an MN-major UMMA operand, no transpose copy —
    other shapes go through one packing pass (``vt_pack_bf16``: zero-padded K-major rows); fp32 operands
    are split into bf16 pieces first (~2^-16 per product, see kernels/bgemm.py).  VT_EXACT_FP32=1 forces the
    FP32-pipe kernel (``vt_gemm_strided``) for A/B comparisons.

    The model's bf16 attention does not come through here (K3 fuses QK^T, softmax and PV); this entry
    point serves callers of the reference API and the unfused path.
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
    scale = float(scale_factor) if apply_scaling else 1.0

    if os.environ.get("VT_EXACT_FP32") == "1":
        step = 32768   # grid.z carries the batch: split very large batches
        for z0 in range(0, batch, step):
            zb = min(step, batch - z0)
            _lib.call("vt_gemm_strided", A[z0:].data_ptr(), B[z0:].data_ptr(), O[z0:].data_ptr(), None,
                      M, N, K, zb, 1,
                      _lib.i64x4(M * K, 0, K, 1), _lib.i64x4(K * N, 0, N, 1), _lib.i64x4(M * N, 0, N, 1),
                      scale, 0, _lib.dtype_code(A), _lib.stream_ptr(A))
        return O

    if A.dtype == torch.bfloat16 and K % 8 == 0 and N % 8 == 0 and A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0:
        # both operands in place: A K-major, B MN-major
        bg.bgemm(A.data_ptr(), B.data_ptr(), O, O.data_ptr(), M, N, K, batch, 1, (M * K, 0, K), (K * N, 0, N),
                 (M * N, 0, N), b_mn=True, scale=scale)
        return O
    pieces = 1 if A.dtype == torch.bfloat16 else bg.split_pieces()
    a = bg.pack(A, A.data_ptr(), M, K, batch, 1, (M * K, 0, K, 1), pieces, pattern=0)
    bt = bg.pack(B, B.data_ptr(), N, K, batch, 1, (K * N, 0, 1, N), pieces, pattern=1)       # B^T rows: K-major
    kk = a.shape[2]
    bg.bgemm(a.data_ptr(), bt.data_ptr(), O, O.data_ptr(), M, N, kk, batch, 1, (M * kk, 0, kk), (N * kk, 0, kk),
             (M * N, 0, N), scale=scale)
    return O
