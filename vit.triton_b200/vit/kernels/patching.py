"""Image -> patch rows (replaces reference vit/kernels/patching.py:54-92)."""
import torch

from . import _lib


def patching(image: torch.Tensor, patch_size: int) -> torch.Tensor:
    """(B, C, H, W) -> (B, N, P*P*C) with each patch row ordered (channel, row, col).

    The model path does not call this (the patch-embedding GEMM gathers pixels itself); it is kept
    because the reference exports it and the exact-fp32 path reuses it as its im2col step.
    """
    assert image.is_cuda, "Image is not on GPU"
    assert len(image.shape) == 4, "The provided matrix for patching should be 4 dimensional (B, C, H, W)"
    B, C, H, W = image.shape
    assert H % patch_size == 0 and W % patch_size == 0, \
        f"Image height and width should be divisible by the patch size, provided: {(H, W)}, {patch_size}"
    image = image.contiguous()
    n = (H // patch_size) * (W // patch_size)
    out = torch.empty((B, n, patch_size * patch_size * C), device=image.device, dtype=image.dtype)
    _lib.call("vt_patching", image.data_ptr(), out.data_ptr(), B, C, H, W, patch_size,
              _lib.dtype_code(image), _lib.stream_ptr(image))
    return out
