"""TEST INFRASTRUCTURE — HuggingFace ``ViTModel`` as the numeric oracle (BASELINE.json north_star).

Random-init models of the named architectures (no network, no pretrained weights), deterministic
from seeds; biases and LayerNorm affines are re-randomised because HF zero/one-initialises them
(modeling_vit.py ``_init_weights``), which would hide bias and affine bugs.
"""
import torch

ARCHS = {
    # name: hidden, layers, heads, mlp, patch, image
    'vit-b16-224': dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, patch_size=16, image_size=224),
    'vit-b16-384': dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, patch_size=16, image_size=384),
    'vit-l16-224': dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, patch_size=16, image_size=224),
    'vit-h14-224': dict(hidden_size=1280, num_hidden_layers=32, num_attention_heads=16, intermediate_size=5120, patch_size=14, image_size=224),
    # small shapes with the same structure (dh 64 / dh 80 + 14-pixel patches) for fast CPU tests
    'tiny-b': dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, patch_size=16, image_size=64),
    'tiny-h': dict(hidden_size=160, num_hidden_layers=2, num_attention_heads=2, intermediate_size=640, patch_size=14, image_size=56),
    # 12 layers x 64-wide heads: the only shape the reference's own loader handles (utils.py:53, load_weights.py:29)
    'tiny-ref': dict(hidden_size=64, num_hidden_layers=12, num_attention_heads=1, intermediate_size=256, patch_size=16, image_size=32),
}


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


def vit_kwargs(arch: str) -> dict:
    """Constructor arguments of vit.vit.VIT for an architecture."""
    a = ARCHS[arch]
    return dict(height=a['image_size'], width=a['image_size'], channels=3, patch_size=a['patch_size'],
                hidden_dim=a['hidden_size'], num_heads=a['num_attention_heads'],
                num_layers=a['num_hidden_layers'], mlp_dim=a['intermediate_size'])


@torch.no_grad()
def hf_forward(model, x: torch.Tensor) -> torch.Tensor:
    return model(pixel_values=x).last_hidden_state
