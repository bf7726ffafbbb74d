"""Attention kernel: isolated timing + in-kernel cycle accounting per slot (developer tool).
VT_LIB=<alternative .so> selects a kernel variant build."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit.kernels import _lib, flash_attention
lib = _lib.load()
lib.vt_debug_set_attn_buffer.argtypes = [ctypes.c_void_p]
lib.vt_debug_set_attn_buffer.restype = None
shapes = ((256, 12, 197, 64), (128, 12, 577, 64))
if len(sys.argv) > 1 and sys.argv[1] == "all":
    shapes += ((128, 16, 197, 64), (64, 16, 257, 80))
for (B, H, N, dh) in shapes:
    qkv = torch.randn(B, N, 3 * H * dh, device="cuda").bfloat16()
    for _ in range(3):
        flash_attention(qkv, H)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        flash_attention(qkv, H)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / 20 * 1e3
    dbg = torch.zeros(3 * 148 * 8, dtype=torch.int64, device="cuda")
    lib.vt_debug_set_attn_buffer(dbg.data_ptr())
    flash_attention(qkv, H)
    torch.cuda.synchronize()
    lib.vt_debug_set_attn_buffer(None)
    nq = (N + 127) // 128
    impl = os.environ.get("VT_ATTN_IMPL", "5")
    print(f"B={B} H={H} N={N} dh={dh}: {us:.1f} us/launch (20 back-to-back)")
    d = dbg[:2 * 148 * 8].view(148, 2, 8).double()
    n = B * H * nq / 296
    names = ["wait-S", "row max", "exchange", "exp", "wait-turn", "norm bound", "-"]
    for g in range(2):
        print(f"   group {g} per item cycles: " + ", ".join(f"{nm} {d[:, g, i].mean()/n:.0f}" for i, nm in enumerate(names))
              + f", total {d[:, g, 7].mean()/n:.0f}")
    m = dbg[2 * 148 * 8:].view(148, 8).double()
    ni = B * H * nq / 148
    print("   MMA issuer per item cycles: " + ", ".join(f"{nm} {m[:, i].mean()/ni:.0f}" for i, nm in
                                                       enumerate(["wait-V", "wait-O-read", "wait-P", "issue PV", "wait Q/K", "issue S"])))
