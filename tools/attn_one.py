"""One attention launch shape for ncu (developer tool): python tools/attn_one.py [B H N dh reps]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit.kernels import flash_attention
B, H, N, dh = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (256, 12, 197, 64)))
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
qkv = torch.randn(B, N, 3 * H * dh, device="cuda").bfloat16()
for _ in range(reps):
    out = flash_attention(qkv, H)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
