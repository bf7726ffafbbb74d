"""Time the fused attention kernel at the C2 / C4 / C5 / C3 shapes (developer tool; VT_LIB selects the build)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit.kernels import flash_attention
for (B, H, N, dh) in ((256, 12, 197, 64), (128, 16, 197, 64), (64, 16, 257, 80), (128, 12, 577, 64)):
    qkv = torch.randn(B, N, 3 * H * dh, device="cuda").bfloat16()
    for _ in range(5):
        out = flash_attention(qkv, H)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(30):
        out = flash_attention(qkv, H)
    e.record(); torch.cuda.synchronize()
    print(f"B={B} H={H} N={N} dh={dh}: {s.elapsed_time(e) / 30 * 1e3:.1f} us per launch")
