"""Fixed cost of one persistent GEMM launch: kernel duration (CUPTI) against tiles per cluster (developer tool)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from vit.kernels import _lib
_lib.load()
K, N = 768, 256
for tiles_per_cluster in (1, 2, 4, 8, 16):
    M = 256 * 74 * tiles_per_cluster
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    def run():
        _lib.call("vt_gemm_bf16", x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, _lib.VT_BF16,
                  bias.data_ptr(), None, N, M, N, K, 0, _lib.stream_ptr(x))
    for _ in range(5): run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10): run()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type.name == "CUDA" and "gemm2" in e.name]
    dur = sorted(e.time_range.elapsed_us() for e in ev)[len(ev) // 2]
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): run()
    e.record(); torch.cuda.synchronize()
    print(f"tiles/cluster {tiles_per_cluster:2d}: kernel {dur:7.1f} us (CUPTI median), back-to-back {s.elapsed_time(e) / 20 * 1e3:7.1f} us per launch")
