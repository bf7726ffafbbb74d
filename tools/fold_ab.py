"""A/B: layernorm_before folded into the QKV GEMM vs the separate LayerNorm kernel, sustained (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit import configs as hf_oracle   # architecture table only
from vit import vit as V
arch = "vit-b16-224"
m = V.VIT(**hf_oracle.vit_kwargs(arch)).to("cuda", torch.bfloat16)
with torch.no_grad():
    for p_ in m.parameters():
        p_.copy_(torch.randn_like(p_) * 0.02)
xs = [torch.randn(256, 3, 224, 224, device="cuda").bfloat16() for _ in range(4)]
def run(n):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        m(xs[i % 4])
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
with torch.no_grad():
    for rep in range(3):
        for fold, mlp in ((False, False), (True, False), (True, True)):
            V.set_layernorm_folding(fold, mlp=mlp)
            run(5)
            print(f"rep {rep} fold={fold} mlp={mlp}: {run(60):.3f} ms/forward", flush=True)
