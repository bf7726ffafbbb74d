"""Time the patch embedding (gather + token-mode GEMM) of the C2 batch (developer tool; VT_LIB selects the build)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit import configs
from vit.vit import VIT
m = VIT(**configs.vit_kwargs("vit-b16-224")).to("cuda", torch.bfloat16)
x = torch.randn(256, 3, 224, 224, device="cuda").bfloat16()
with torch.no_grad():
    for _ in range(5):
        y = m.embeddings(x)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        y = m.embeddings(x)
    e.record(); torch.cuda.synchronize()
print(f"patch embedding (gather + GEMM), 256 x 224^2 bf16: {s.elapsed_time(e) / 50 * 1e3:.1f} us; checksum {y.float().abs().sum().item():.6e}")
