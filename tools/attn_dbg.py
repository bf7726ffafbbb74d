"""In-kernel cycle accounting of the persistent attention kernel (developer tool)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit.kernels import _lib, flash_attention
lib = _lib.load()
lib.vt_debug_set_attn_buffer.argtypes = [ctypes.c_void_p]
lib.vt_debug_set_attn_buffer.restype = None
for (B, H, N) in ((256, 12, 197), (128, 12, 577)):
    qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16()
    for _ in range(3):
        flash_attention(qkv, H)
    dbg = torch.zeros(2 * 148 * 8, dtype=torch.int64, device="cuda")
    lib.vt_debug_set_attn_buffer(dbg.data_ptr())
    flash_attention(qkv, H)
    torch.cuda.synchronize()
    lib.vt_debug_set_attn_buffer(None)
    d = dbg.view(296, 8).double()
    n = d[:, 4].mean()
    print(f"B={B} N={N}: items/slot {n:.1f}; per item cycles: wait-S {d[:,0].mean()/n:.0f}, softmax {d[:,1].mean()/n:.0f}, "
          f"(pass1 {d[:,6].mean()/n:.0f}) wait-O {d[:,2].mean()/n:.0f}, epilogue {d[:,3].mean()/n:.0f}, total {d[:,5].mean()/n:.0f} (kernel cycles {d[:,5].mean():.0f})")
