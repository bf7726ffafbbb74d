"""Run one C2-shaped GEMM launch sequence (for ncu): python tools/prof_gemm.py K N gelu res"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit.kernels import _lib
K, N, act, res = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
M = 256 * 197
x = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device="cuda")
r = torch.randn(M, N, device="cuda").bfloat16() if res else None
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(6):
    _lib.call("vt_gemm_bf16", x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, _lib.VT_BF16,
              bias.data_ptr(), None if r is None else r.data_ptr(), N, M, N, K, act, _lib.stream_ptr(x))
torch.cuda.synchronize()
print("ok")
