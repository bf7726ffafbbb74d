"""In-kernel cycle accounting of the GEMM launches INSIDE a C2 forward (developer tool): the same counters as
tools/gemm_dbg.py, but taken while the model runs (cold operands, the power-capped clock of the real step)."""
import collections, ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit import configs
from vit.vit import VIT
from vit.kernels import _lib
lib = _lib.load()
lib.vt_debug_set_buffer.argtypes = [ctypes.c_void_p]
lib.vt_debug_set_buffer.restype = None
arch = sys.argv[1] if len(sys.argv) > 1 else "vit-b16-224"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = VIT(**configs.vit_kwargs(arch)).to("cuda", torch.bfloat16)
with torch.no_grad():
    for p_ in m.parameters():
        p_.copy_(torch.randn_like(p_) * 0.02)
S = configs.ARCHS[arch]["image_size"]
x = torch.randn(B, 3, S, S, device="cuda").bfloat16()
bufs = []
def hook(name, before, args=None):
    if name not in ("vt_gemm_bf16_ln", "vt_gemm_bf16"):
        return
    if before:
        b = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
        if name == "vt_gemm_bf16_ln":
            M, N, K, gelu, res = args[9], args[10], args[11], args[12], args[7]
        else:
            M, N, K, gelu, res = args[10], args[11], args[12], args[13], args[8]
        bufs.append(((M, N, K, int(gelu), res is not None), b))
        lib.vt_debug_set_buffer(b.data_ptr())
    else:
        lib.vt_debug_set_buffer(None)
with torch.no_grad():
    for _ in range(30):
        m(x)
    torch.cuda.synchronize()
    _lib.event_hook = hook
    for _ in range(3):
        m(x)
    _lib.event_hook = None
    torch.cuda.synchronize()
from vit.utils import capture_cuda_graph
# the same counters inside a CUDA-graph replay: the debug buffers are allocated (and their pointers baked into the
# launches) at capture time, the last replay leaves its timestamps in them
eager_bufs = list(bufs)
bufs.clear()
with torch.no_grad():
    _lib.event_hook = hook
    g, _ = capture_cuda_graph(m, x)
    _lib.event_hook = None
    for _ in range(20):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"graph replay: {e0.elapsed_time(e1) / 50:.3f} ms per forward")
agg = collections.OrderedDict()
for key, b in eager_bufs:
    agg.setdefault(key, []).append(b.view(148, 8).double())
# launch skew and gaps from the absolute in-kernel timestamps of the last graph replay (48 GEMM launches)
last = bufs[-48:]
torch.cuda.synchronize()
spans, skews, gaps = [], [], []
prev_end = None
for key, b in last:
    d = b.view(148, 8)
    st, en = d[:, 6], d[:, 6] + d[:, 7]
    spans.append((en.max() - st.min()).item() / 1e3)
    skews.append(((st.max() - st.min()).item() / 1e3, (en.max() - en.min()).item() / 1e3))
    if prev_end is not None:
        gaps.append((key, (st.min() - prev_end).item() / 1e3))
    prev_end = en.max()
print(f"GEMM launches of one forward: sum of (last CTA end - first CTA start) {sum(spans)/1e3:.3f} ms; "
      f"mean start skew {sum(a for a, _ in skews)/len(skews):.1f} us, mean end skew {sum(b for _, b in skews)/len(skews):.1f} us")
print("time from the end of a GEMM to the start of the next GEMM (us; attention sits inside the QKV -> out-proj gap):")
print("  " + ", ".join(f"N={k[1]},K={k[2]}:{g:.1f}" for k, g in gaps[:8]))
tot_ns = 0.0
for (M, N, K, gelu, res), ds in agg.items():
    tot_ns += torch.stack(ds).mean(0)[:, 7].mean().item() * len(ds) / 3
print(f"sum of the in-kernel wall times of the GEMM launches of one forward: {tot_ns / 1e6:.3f} ms")
for (M, N, K, gelu, res), ds in agg.items():
    d = torch.stack(ds).mean(0)
    tiles = -(-M // 256) * -(-N // 256)
    per = tiles / 74
    lead = d[0::2]
    print(f"M={M} N={N} K={K} gelu={gelu} res={res} x{len(ds)}: tiles/cluster {per:.1f}; total cyc {d[:,5].mean():.0f} (max {d[:,5].max():.0f}); per tile: "
          f"epi wait-tfull {d[:,0].mean()/per:.0f}, epi busy {d[:,1].mean()/per:.0f}, store-drain {d[:,6].mean()/per:.0f}, "
          f"mma wait-full {lead[:,2].mean()/per:.0f}, mma wait-tempty {lead[:,3].mean()/per:.0f}, prod wait-empty {d[:,4].mean()/per:.0f}, "
          f"period {d[:,5].mean()/per:.0f}; in-kernel wall {d[:,7].mean()/1e3:.1f} us = SM clock {d[:,5].mean()/d[:,7].mean():.3f} GHz")

# is the finishing spread tied to particular CTAs (SMs) or random?  per-CTA duration relative to the launch mean,
# over the launches of one shape in the last graph replay
import collections as _c
by_shape = _c.OrderedDict()
for key, b in last:
    by_shape.setdefault(key, []).append(b.view(148, 8)[:, 7].double())
for key, ws in by_shape.items():
    w = torch.stack(ws)                      # [launches, 148] wall ns
    rel = w / w.mean(1, keepdim=True)
    per_cta = rel.mean(0)
    grp = per_cta.view(37, 4).mean(1)
    print(f"N={key[1]} K={key[2]}: per-CTA mean relative duration min {per_cta.min():.3f} max {per_cta.max():.3f}; "
          f"std of the per-CTA means {per_cta.std():.4f}, mean std within a CTA over launches {rel.std(0).mean():.4f}; "
          f"slowest groups {[int(i) for i in grp.argsort(descending=True)[:5]]} ({grp.max():.3f}), "
          f"fastest {[int(i) for i in grp.argsort()[:5]]} ({grp.min():.3f})")
