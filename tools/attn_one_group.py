import ctypes, os, sys
sys.path.insert(0, "/root/repo/vit.triton_b200")
import torch
from vit.kernels import _lib, flash_attention
lib = _lib.load()
lib.vt_debug_set_attn_buffer.argtypes = [ctypes.c_void_p]
lib.vt_debug_set_attn_buffer.restype = None
for (B, H, N) in ((6, 12, 197), (12, 12, 197), (256, 12, 197)):
    qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16()
    for _ in range(3): flash_attention(qkv, H)
    dbg = torch.zeros(3 * 148 * 8, dtype=torch.int64, device="cuda")
    lib.vt_debug_set_attn_buffer(dbg.data_ptr()); flash_attention(qkv, H); torch.cuda.synchronize(); lib.vt_debug_set_attn_buffer(None)
    mma = dbg[2 * 148 * 8:].view(148, 8).double()
    d = dbg[:2 * 148 * 8].view(148, 2, 8).double()
    items = B * H * 2
    names = ["wait-S", "pass1", "sync", "pass2", "wait-O", "O-read", "epi"]
    print(f"B={B}: MMA issuer per item: wait-V {mma[:,0].mean()/max(1,items/148):.0f}, wait-O-read {mma[:,1].mean()/max(1,items/148):.0f}, wait-P {mma[:,2].mean()/max(1,items/148):.0f}, issue PV + next S {mma[:,3].mean()/max(1,items/148):.0f}")
    for g in range(2):
        n = max(1.0, (items / 148 + (1 - g)) // 2) if items < 148*2 else items / 296
        act = d[:, g, 7] > 0
        if act.sum() == 0: continue
        print(f"B={B}: group {g}: active CTAs {int(act.sum())}: " + ", ".join(f"{nm} {d[act][:, g, i].mean()/n:.0f}" for i, nm in enumerate(names)) + f" total {d[act][:, g, 7].mean()/n:.0f} (per item, n={n})")
