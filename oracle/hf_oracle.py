"""TEST INFRASTRUCTURE — HuggingFace ``ViTModel`` as the numeric oracle (BASELINE.json north_star).

Random-init models of the named architectures (no network, no pretrained weights), deterministic
from seeds; biases and LayerNorm affines are re-randomised because HF zero/one-initialises them
(modeling_vit.py ``_init_weights``), which would hide bias and affine bugs.
"""
import torch

import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit.triton_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from vit.configs import ARCHS, vit_kwargs  # noqa: E402,F401  (one table for the product and its checker)


def build_hf(arch: str, seed: int = 0, randomize_affine: bool = True):
    from transformers import ViTConfig, ViTModel
    cfg = ViTConfig(**ARCHS[arch])
    torch.manual_seed(seed)
    model = ViTModel(cfg, add_pooling_layer=False).eval()
    if randomize_affine:
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if name.endswith('.bias'):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
                elif 'layernorm' in name and name.endswith('.weight'):
                    p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.02)
    return model


def make_input(arch: str, batch: int, seed: int = 1234) -> torch.Tensor:
    s = ARCHS[arch]['image_size']
    g = torch.Generator().manual_seed(seed)
    return torch.randn((batch, 3, s, s), generator=g)


@torch.no_grad()
def hf_forward(model, x: torch.Tensor) -> torch.Tensor:
    return model(pixel_values=x).last_hidden_state
