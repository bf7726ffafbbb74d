"""Does the power-capped step gain from running attention NEXT TO the GEMMs (disjoint SM sets, two streams, two half
batches in a software pipeline) instead of after them?  Developer experiment: the kernels of 12 layers at the C2
shapes, (a) as the model runs them today (full batch, one stream, every kernel on all SMs), (b) two half batches,
GEMMs on `gemm_groups` x 4 SMs in one stream, attention on the other SMs in a second stream, ordered
QKV_A | proj/fc1/fc2_B | QKV_B | proj/fc1/fc2_A with attention_X between QKV_X and proj_X."""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit.kernels import _lib, flash_attention
lib = _lib.load()
lib.vt_debug_set_sm_partition.argtypes = [ctypes.c_int, ctypes.c_int]
lib.vt_debug_set_sm_partition.restype = None
B, N, D, F, H, L = 256, 197, 768, 3072, 12, 12
dev = "cuda"
def mk(Bh):
    # every kernel reads FIXED random inputs and writes into separate outputs: data that feeds back through 12 layers
    # without LayerNorm overflows to inf / NaN, and constant bit patterns draw so little power that the loop runs 15 %
    # faster than the model (7.05 vs 8.46 ms on one box) — the step time depends on the data under a power cap
    M = Bh * N
    r = lambda *shape: torch.randn(*shape, device=dev).bfloat16()
    e = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.bfloat16)
    return {"x": r(M, D), "qkv_in": r(Bh, N, 3 * D), "ctx_in": r(M, D), "x2_in": r(M, D), "mid_in": r(M, F),
            "qkv_out": e(M, 3 * D), "x2_out": e(M, D), "mid_out": e(M, F), "x_out": e(M, D), "M": M, "B": Bh}
W = {"qkv": (torch.randn(3 * D, D, device=dev) / math.sqrt(D)).bfloat16(), "proj": (torch.randn(D, D, device=dev) / math.sqrt(D)).bfloat16(),
     "fc1": (torch.randn(F, D, device=dev) / math.sqrt(D)).bfloat16(), "fc2": (torch.randn(D, F, device=dev) / math.sqrt(F)).bfloat16()}
bias = {k: torch.randn(v.shape[0], device=dev) for k, v in W.items()}
def gemm(a, w, b, out, M, Nn, K, gelu=0, res=None):
    _lib.call("vt_gemm_bf16", a.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), Nn, _lib.VT_BF16, b.data_ptr(),
              None if res is None else res.data_ptr(), Nn, M, Nn, K, gelu, _lib.stream_ptr(a))
def qkv(t): gemm(t["x"], W["qkv"], bias["qkv"], t["qkv_out"], t["M"], 3 * D, D)
def attn(t): t["ctx"] = flash_attention(t["qkv_in"], H)
def mlp(t):
    gemm(t["ctx_in"], W["proj"], bias["proj"], t["x2_out"], t["M"], D, D, 0, t["x"])
    gemm(t["x2_in"], W["fc1"], bias["fc1"], t["mid_out"], t["M"], F, D, 1)
    gemm(t["mid_in"], W["fc2"], bias["fc2"], t["x_out"], t["M"], D, F, 0, t["x2_in"])

def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

full = mk(B)
def serial():
    for _ in range(L):
        qkv(full); attn(full); mlp(full)
lib.vt_debug_set_sm_partition(0, 0)
g = torch.cuda.CUDAGraph()
serial(); torch.cuda.synchronize()
with torch.cuda.graph(g):
    serial()
print(f"(a) full batch, one stream, all SMs: {timed(g.replay, 60):.3f} ms per {L} layers", flush=True)

halves = [mk(B // 2), mk(B // 2)]
for gg in (31, 30, 28):
    na = 148 - 4 * gg
    def piped():
        main = torch.cuda.current_stream()
        side = piped.side
        evq = [None, None]
        # prologue: QKV of both halves
        for l in range(L):
            for h in (0, 1):
                lib.vt_debug_set_sm_partition(gg, na)
                qkv(halves[h])
                ev = torch.cuda.Event(); ev.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    attn(halves[h])
                    eva = torch.cuda.Event(); eva.record(side)
                evq[h] = eva
                # the MLP GEMMs of the OTHER half (its attention was launched one slot earlier)
                o = 1 - h
                if evq[o] is not None and (l > 0 or h == 1):
                    main.wait_event(evq[o])
                    mlp(halves[o])
                    evq[o] = None
        for h in (0, 1):
            if evq[h] is not None:
                main.wait_event(evq[h]); mlp(halves[h]); evq[h] = None
    piped.side = torch.cuda.Stream()
    piped(); torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        piped()
    lib.vt_debug_set_sm_partition(0, 0)
    print(f"(b) two half batches, GEMMs on {4 * gg} SMs + attention on {na} SMs: {timed(g2.replay, 60):.3f} ms per {L} layers", flush=True)
# (c) the same pipeline without partitioning (every kernel asks for all SMs)
