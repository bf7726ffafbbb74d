"""Eager vs CUDA-graph forward timing at C2 + per-kernel breakdown via CUDA events (developer tool)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit import configs as hf_oracle   # architecture table only
from vit.vit import VIT
from vit.kernels import _lib
from vit.utils import capture_cuda_graph
arch = sys.argv[1] if len(sys.argv) > 1 else "vit-b16-224"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = VIT(**hf_oracle.vit_kwargs(arch)).to("cuda", torch.bfloat16)
with torch.no_grad():
    for p_ in m.parameters():
        p_.copy_(torch.randn_like(p_) * 0.02)
S = hf_oracle.ARCHS[arch]["image_size"]
x = torch.randn(B, 3, S, S, device="cuda").bfloat16()
def timeit(fn, n=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
with torch.no_grad():
    eager = timeit(lambda: m(x))
    g, out = capture_cuda_graph(m, x)
    graph = timeit(lambda: g.replay())
    print(f"{arch} b{B}: eager {eager:.3f} ms, graph {graph:.3f} ms")
    # per-kernel breakdown
    evs = []
    def hook(name, before, args=None):
        ev = torch.cuda.Event(enable_timing=True); ev.record(); evs.append((name, before, ev))
    _lib.event_hook = hook
    m(x)
    _lib.event_hook = None
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    tot = 0.0
    for i in range(0, len(evs), 2):
        name = evs[i][0]; t = evs[i][2].elapsed_time(evs[i + 1][2]); tot += t
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    for k, (n, t) in agg.items():
        print(f"  {k:18s} x{n:3d}  {t*1e3:9.1f} us total  {t/n*1e3:8.1f} us each")
    print(f"  sum of kernels {tot:.3f} ms; first..last event {evs[0][2].elapsed_time(evs[-1][2]):.3f} ms")

# CUPTI view (accurate kernel durations and idle gaps)
try:
    from torch.profiler import profile, ProfilerActivity
    with torch.no_grad():
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                m(x)
            torch.cuda.synchronize()
    evts = [e for e in prof.events() if e.device_type.name == "CUDA"]
    evts.sort(key=lambda e: e.time_range.start)
    agg = collections.OrderedDict()
    for e in evts:
        a = agg.setdefault(e.name[:60], [0, 0.0]); a[0] += 1; a[1] += e.time_range.elapsed_us()
    tot = sum(t for _, t in agg.values())
    span = evts[-1].time_range.end - evts[0].time_range.start
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  [cupti] {t/3:9.1f} us/fwd  x{n//3:3d}  {t/n:8.1f} us each  {k}")
    print(f"  [cupti] kernels {tot/3/1e3:.3f} ms/fwd, span {span/3/1e3:.3f} ms/fwd, idle {(span-tot)/3/1e3:.3f} ms/fwd")
except Exception as ex:
    print("profiler failed:", ex)
