"""Stride == kernel convolution — host entry points replacing reference vit/kernels/conv2d.py:100-167."""
from typing import Tuple

import torch

from . import _lib


def conv2d(input: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """(B, C, H, W) * (O, C, kh, kw) + bias -> (B, O, H/kh, W/kw); stride = kernel size, no padding.

    Checks and messages follow ``conv2d_triton`` (conv2d.py:105-114).  Output dtype and device follow
    the input (the reference hard-codes fp32 on cuda:0, conv2d.py:6-7,116).  Inside ``Embeddings`` the
    model does not call this: the patch-embedding GEMM (K2) also adds CLS and position embeddings.
    """
    assert input.is_cuda and kernel.is_cuda, 'Input or kernel is not on GPU'
    assert len(input.shape) == 4, f'Input needs to be 4 dimensional, provided: {input.shape}'
    assert len(kernel.shape) == 4, f'Kernel size needs to be 4 dimensional, provided: {kernel.shape}'
    assert bias.shape[0] == kernel.shape[0], 'Bias dimension should be same as the kernel 1st dimension'
    batch_size, channels, height, width = input.shape
    num_kernels, kernel_depth, kernel_height, kernel_width = kernel.shape
    assert height % kernel_height == 0 and width % kernel_width == 0, \
        "Input height and width should be divisible by the kernel height and width"
    assert channels == kernel_depth, \
        f"Kernel channel depth ({kernel_depth}) and input channel depth ({channels}) should be same"
    assert input.dtype == kernel.dtype == bias.dtype, \
        f"Input dtypes need to be same, provided {input.dtype}, {kernel.dtype}, {bias.dtype}"

    input = input.contiguous()
    kernel = kernel.contiguous()
    bias = bias.contiguous()
    oh, ow = height // kernel_height, width // kernel_width
    output = torch.empty((batch_size, num_kernels, oh, ow), device=input.device, dtype=input.dtype)
    if output.numel() == 0:
        return output
    code = _lib.dtype_code(input)
    stream = _lib.stream_ptr(input)

    if kernel_height == kernel_width and batch_size <= 32768:
        # im2col rows + strided GEMM writing NCHW directly: out[b, o, p] = patches[b, p, :] . w[o, :]
        from .patching import patching
        P = kernel_height
        K = channels * P * P
        n = oh * ow
        patches = patching(input, P)
        _lib.call("vt_gemm_strided", patches.data_ptr(), kernel.data_ptr(), output.data_ptr(),
                  bias.data_ptr(), n, num_kernels, K, batch_size, 1,
                  _lib.i64x4(n * K, 0, K, 1), _lib.i64x4(0, 0, 1, K),
                  _lib.i64x4(num_kernels * n, 0, 1, n), 1.0, 0, code, stream)
        return output

    _lib.call("vt_conv2d", input.data_ptr(), kernel.data_ptr(), bias.data_ptr(), output.data_ptr(),
              batch_size, channels, height, width, num_kernels, kernel_height, kernel_width, code, stream)
    return output


class Conv2DTriton(torch.nn.Module):
    """Drop-in for the reference module of the same name (conv2d.py:153-167): same constructor,
    parameter names and shapes ((O, I, kh, kw) weight, (O,) bias)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: Tuple):
        super().__init__()
        assert type(kernel_size) == tuple and len(kernel_size) == 2, 'Param kernel size should be a tuple of size 2'
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.weight = torch.nn.Parameter(
            torch.zeros(self.out_channels, self.in_channels, self.kernel_size[0], self.kernel_size[1]))
        self.bias = torch.nn.Parameter(torch.zeros(self.out_channels))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return conv2d(x, self.weight, self.bias)


Conv2DB200 = Conv2DTriton
