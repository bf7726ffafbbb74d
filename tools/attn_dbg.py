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
    nq = (N + 127) // 128
    n = B * H * nq / 296
    names = ["wait-S", "pass1", "max-sync", "pass2", "wait-O", "O-read", "epilogue"]
    print(f"B={B} N={N}: items/slot {n:.1f}; per item cycles: " + ", ".join(f"{nm} {d[:,i].mean()/n:.0f}" for i, nm in enumerate(names))
          + f", total {d[:,7].mean()/n:.0f}")
