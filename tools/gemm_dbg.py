"""In-kernel cycle accounting of the 2-CTA GEMM (developer tool): who waits for whom."""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit.kernels import _lib
lib = _lib.load()
lib.vt_debug_set_buffer.argtypes = [ctypes.c_void_p]
lib.vt_debug_set_buffer.restype = None
M = int(os.environ.get("VT_DBG_M", 256 * 197))
dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
CASES = ((768, 2304, 0, False, ""), (768, 2304, 0, False, "lnf"), (768, 2304, 0, False, "lnz"), (768, 3072, 1, False, ""),
         (768, 3072, 1, False, "lnf"), (768, 3072, 1, False, "lnz"),
         (768, 3072, 0, False, ""), (768, 768, 0, True, ""), (768, 768, 0, True, "stats"), (3072, 768, 0, True, ""),
         (3072, 768, 0, True, "stats"))
for (K, N, act, res, mode) in CASES:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").bfloat16() if res else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    rowstats = torch.randn(M, K // 128, 2, device="cuda").abs() + 1.0
    colsum = torch.randn(N, device="cuda")
    stats_out = torch.empty(M, N // 128, 2, device="cuda")
    def run():
        if mode in ("lnf", "lnz"):
            _lib.call("vt_gemm_bf16_ln", x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, bias.data_ptr(), None, 0,
                      M, N, K, act, rowstats.data_ptr(), colsum.data_ptr() if mode == "lnf" else None, K, 1e-12, None,
                      _lib.stream_ptr(x))
        elif mode == "stats":
            _lib.call("vt_gemm_bf16_ln", x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, bias.data_ptr(), r.data_ptr(), N,
                      M, N, K, 0, None, None, 0, 0.0, stats_out.data_ptr(), _lib.stream_ptr(x))
        else:
            _lib.call("vt_gemm_bf16", x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, _lib.VT_BF16,
                      bias.data_ptr(), None if r is None else r.data_ptr(), N, M, N, K, act, _lib.stream_ptr(x))
    lib.vt_debug_set_buffer(None)
    for _ in range(5):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    dbg.zero_()
    lib.vt_debug_set_buffer(dbg.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.vt_debug_set_buffer(None)
    d = dbg.view(148, 8).double()
    tiles = math.ceil(M / 256) * math.ceil(N / 256)
    per_cta = tiles / 74
    lead = d[0::2]
    print(f"K={K} N={N} gelu={act} res={res} {mode} [{us:.1f} us back-to-back]: tiles/cluster {per_cta:.1f}; total cyc {d[:,5].mean():.0f}; per tile: "
          f"epi wait-tfull {d[:,0].mean()/per_cta:.0f}, epi busy {d[:,1].mean()/per_cta:.0f}, epi store-drain {d[:,6].mean()/per_cta:.0f}, "
          f"mma wait-full {lead[:,2].mean()/per_cta:.0f}, mma wait-tempty {lead[:,3].mean()/per_cta:.0f}, prod wait-empty {d[:,4].mean()/per_cta:.0f}, "
          f"period {d[:,5].mean()/per_cta:.0f}")
