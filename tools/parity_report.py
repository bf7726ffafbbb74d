import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/vit.triton_b200")
import torch
from oracle import hf_oracle
from vit.utils import transfer_pretrained_weights
from vit.vit import VIT
def cos(a, b): return torch.nn.functional.cosine_similarity(a.flatten().float(), b.flatten().float(), dim=0).item()
for arch, dtype, nimg in (("vit-b16-224", torch.bfloat16, 8), ("vit-b16-224", torch.float32, 1), ("vit-b16-384", torch.bfloat16, 2), ("vit-l16-224", torch.bfloat16, 2), ("vit-h14-224", torch.bfloat16, 2)):
    hf = hf_oracle.build_hf(arch, seed=0)
    m = VIT(**hf_oracle.vit_kwargs(arch)); transfer_pretrained_weights(hf, m, verbose=False)
    m = m.to("cuda", dtype).eval()
    x = hf_oracle.make_input(arch, nimg)
    want = hf_oracle.hf_forward(hf, x)
    with torch.no_grad(): got = m(x.to("cuda", dtype)).float().cpu()
    print(f"{arch} {str(dtype)[6:]} x{nimg}: cosine {cos(got, want):.6f}  max-abs {(got - want).abs().max().item():.3e}  |ref|max {want.abs().max().item():.2f}")
