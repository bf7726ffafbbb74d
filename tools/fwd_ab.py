"""Sustained C2 forward time of the current settings (developer tool; run several times with different
VT_* environment variables in ONE gpurun call to A/B them on the same box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit import configs
from vit import vit as V
arch = sys.argv[1] if len(sys.argv) > 1 else "vit-b16-224"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = V.VIT(**configs.vit_kwargs(arch)).to("cuda", torch.bfloat16)
with torch.no_grad():
    for p_ in m.parameters():
        p_.copy_(torch.randn_like(p_) * 0.02)
S = configs.ARCHS[arch]["image_size"]
xs = [torch.randn(batch, 3, S, S, device="cuda").bfloat16() for _ in range(4)]
def run(n):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        m(xs[i % 4])
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
from vit.utils import capture_cuda_graph
with torch.no_grad():
    run(10)
    eager = run(100)
    g, out = capture_cuda_graph(m, xs[0])
    def replay(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n):
            g.replay()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / n
    replay(10)
    graph = replay(100)
    ref = m(xs[0])
    g.replay(); torch.cuda.synchronize()
    same = torch.equal(ref, out)
    print(" ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("VT_")) or "defaults",
          f": eager {eager:.3f} ms/forward, graph replay {graph:.3f} ms/forward (replay == eager: {same})", flush=True)
