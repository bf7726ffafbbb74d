"""Worker of test_launch_modes_are_bit_identical (tests/test_gpu_model.py): one ViT-B/16 bf16 forward of 24 images plus
one GEMM of every epilogue at shapes with odd tile counts, under whatever VT_GEMM_QUAD / VT_PDL / VT_ATTN_NO_BOUND the
environment sets (the library reads them once per process); the results go to the file named on the command line."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vit.triton_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch
from vit import configs
from vit.kernels import _lib
from vit.utils import capture_cuda_graph
from vit.vit import VIT

torch.manual_seed(7)
dev = "cuda"
model = VIT(**configs.vit_kwargs("vit-b16-224"))
with torch.no_grad():
    for p_ in model.parameters():
        p_.copy_(torch.randn_like(p_) * 0.02)
model = model.to(dev, torch.bfloat16)
x = torch.randn(24, 3, 224, 224).to(dev, torch.bfloat16)
out = {}
with torch.no_grad():
    out["eager"] = model(x).clone()
    graph, static_out = capture_cuda_graph(model, x)
    graph.replay()
    torch.cuda.synchronize()
    out["graph"] = static_out.clone()
for (M, K, N, gelu, res) in ((197 * 7, 768, 768, 0, True), (513, 1280, 1280, 1, False), (197, 768, 2304, 0, False),
                             (2000, 3072, 768, 0, True)):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).bfloat16() if res else None
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):     # back to back: dependent launches
        _lib.call("vt_gemm_bf16", a.data_ptr(), K, w.data_ptr(), K, o.data_ptr(), N, _lib.VT_BF16, bias.data_ptr(),
                  None if r is None else r.data_ptr(), N, M, N, K, gelu, _lib.stream_ptr(a))
    torch.cuda.synchronize()
    out[f"gemm_{M}_{K}_{N}_{gelu}_{int(res)}"] = o.clone()
torch.save({k: v.cpu() for k, v in out.items()}, sys.argv[1])
