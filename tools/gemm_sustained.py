"""Sustained (power-capped) throughput of the 2-CTA GEMM launches of one ViT-B layer, back to back for ~1 s per
shape (developer tool; VT_LIB selects an alternative build)."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit.kernels import _lib
_lib.load()
M = int(os.environ.get("VT_DBG_M", 256 * 197))
CASES = ((768, 2304, 0, False), (768, 3072, 1, False), (768, 768, 0, True), (3072, 768, 0, True))
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
for (K, N, act, res) in CASES:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").bfloat16() if res else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    def run():
        _lib.call("vt_gemm_bf16", x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, _lib.VT_BF16,
                  bias.data_ptr(), None if r is None else r.data_ptr(), N, M, N, K, act, _lib.stream_ptr(x))
    for _ in range(20):
        run()
    torch.cuda.synchronize()
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    n = max(50, int(secs * 1e6 / us))
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"K={K} N={N} gelu={act} res={res}: {us:.1f} us per launch sustained over {n} launches = {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")
    if os.environ.get("VT_CUBLAS"):
        # library comparison on the same shape: cuBLASLt through torch (bias epilogue only; no GELU / residual)
        wb = bias.bfloat16()
        def run_lib():
            torch.nn.functional.linear(x, w, wb, ) if False else torch.addmm(wb, x, w.t(), out=out)
        for _ in range(20):
            run_lib()
        torch.cuda.synchronize()
        n2 = max(50, int(secs * 1e6 / us))
        e0.record()
        for _ in range(n2):
            run_lib()
        e1.record()
        torch.cuda.synchronize()
        us2 = e0.elapsed_time(e1) / n2 * 1e3
        print(f"    cuBLASLt addmm (bias only) on the same shape: {us2:.1f} us = {2.0 * M * N * K / us2 / 1e6:.0f} TFLOP/s")
