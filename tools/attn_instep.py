"""In-kernel cycle accounting of the attention launches INSIDE a C2 forward (developer tool; needs the
-DVT_ATTN5_DBG build: VT_LIB=.../libvitb200_attndbg.so)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit.triton_b200"))
import torch
from vit import configs
from vit.vit import VIT
from vit.kernels import _lib
lib = _lib.load()
lib.vt_debug_set_attn_buffer.argtypes = [ctypes.c_void_p]
lib.vt_debug_set_attn_buffer.restype = None
arch, B = "vit-b16-224", 256
m = VIT(**configs.vit_kwargs(arch)).to("cuda", torch.bfloat16)
with torch.no_grad():
    for p_ in m.parameters():
        p_.copy_(torch.randn_like(p_) * 0.02)
x = torch.randn(B, 3, 224, 224, device="cuda").bfloat16()
bufs = []
def hook(name, before, args=None):
    if name != "vt_flash_attn":
        return
    if before:
        b = torch.zeros(3 * 148 * 8, dtype=torch.int64, device="cuda")
        bufs.append(b)
        lib.vt_debug_set_attn_buffer(b.data_ptr())
    else:
        lib.vt_debug_set_attn_buffer(None)
with torch.no_grad():
    for _ in range(30):
        m(x)
    torch.cuda.synchronize()
    _lib.event_hook = hook
    for _ in range(3):
        m(x)
    _lib.event_hook = None
    torch.cuda.synchronize()
dbg = torch.stack(bufs).double().mean(0)
H, N = 12, 197
nq = 2
d = dbg[:2 * 148 * 8].view(148, 2, 8)
n = B * H * nq / 296
names = ["wait-S", "row max", "exchange", "exp", "wait-turn", "norm bound", "-"]
for g in range(2):
    print(f"   group {g} per item cycles: " + ", ".join(f"{nm} {d[:, g, i].mean()/n:.0f}" for i, nm in enumerate(names))
          + f", total {d[:, g, 7].mean()/n:.0f}")
mm = dbg[2 * 148 * 8:].view(148, 8)
ni = B * H * nq / 148
print("   MMA issuer per item cycles: " + ", ".join(f"{nm} {mm[:, i].mean()/ni:.0f}" for i, nm in
                                                   enumerate(["wait-V", "wait-O-read", "wait-P", "issue PV", "wait Q/K", "issue S"])))
