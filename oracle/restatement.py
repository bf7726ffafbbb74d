"""TEST INFRASTRUCTURE — plain-PyTorch fp32 CPU restatement of cmeraki/vit.triton's forward pass.

It consumes the *custom* state-dict (the reference's per-head parameter layout), so it checks the
loader and the weight layouts as well as the arithmetic.  Every function cites the reference lines
it follows.  Pinning: the reference's own tests hold no golden vectors and its Triton
kernels cannot execute without a GPU, so this restatement is pinned against HuggingFace
``ViTModel`` fp32 — the comparison target the reference itself uses (vit/vit.py:273,
utils.py:136-178) and the oracle BASELINE.json names — by tests/test_oracle.py (<= 2e-5 max-abs),
and against the committed fixtures in tests/golden/ produced by oracle/make_golden.py.
"""
import math
import re
from typing import Dict

import torch


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    # vit/kernels/activations.py:19-20 — 0.5 * x * (1 + erf(x / sqrt(2)))
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def matmul(a: torch.Tensor, w: torch.Tensor, bias=None, activation=None) -> torch.Tensor:
    # vit/kernels/matmul.py:73-108 — fp32 accumulate, + bias, optional exact GELU
    out = a @ w
    if bias is not None:
        out = out + bias
    if activation == 'gelu':
        out = gelu_erf(out)
    return out


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    # vit/kernels/layernorm.py:51-85 — mean; biased variance of centred values; sqrt(var + eps)
    mean = x.mean(dim=-1, keepdim=True)
    centred = x - mean
    var = (centred * centred).mean(dim=-1, keepdim=True)
    return w * (centred / torch.sqrt(var + eps)) + b


def softmax(x: torch.Tensor) -> torch.Tensor:
    # vit/kernels/softmax.py:26-31 — subtract row max, exp, divide by row sum
    e = torch.exp(x - x.max(dim=-1, keepdim=True).values)
    return e / e.sum(dim=-1, keepdim=True)


def conv2d_patches(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    # vit/kernels/conv2d.py:19-97 — stride == kernel, no padding: per-patch dot products
    B, C, H, W = x.shape
    O, _, kh, kw = w.shape
    p = x.reshape(B, C, H // kh, kh, W // kw, kw).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // kh) * (W // kw), C * kh * kw)
    out = p @ w.reshape(O, -1).t() + b          # (B, n, O)
    return out.transpose(1, 2).reshape(B, O, H // kh, W // kw)


def infer_hparams(sd: Dict[str, torch.Tensor]) -> dict:
    layers = 1 + max(int(m.group(1)) for m in (re.match(r"encoder\.layer\.(\d+)\.", k) for k in sd) if m)
    heads = 1 + max(int(m.group(1)) for m in
                    (re.match(r"encoder\.layer\.0\.attention\.attention\.(\d+)\.", k) for k in sd) if m)
    return dict(layers=layers, heads=heads, hidden=sd['layernorm.weight'].numel(),
                patch=sd['embeddings.projection.weight'].shape[-1])


def vit_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, return_layers: bool = False):
    """VIT.forward of the reference (vit/vit.py:240-247) on a custom state-dict, fp32 on CPU."""
    hp = infer_hparams(sd)
    sd = {k: v.float() for k, v in sd.items()}
    x = x.float()

    # Embeddings.forward — vit/vit.py:188-200
    t = conv2d_patches(x, sd['embeddings.projection.weight'], sd['embeddings.projection.bias'])
    t = t.flatten(2).transpose(1, 2)
    cls = sd['embeddings.cls_token'].expand(t.shape[0], -1, -1)
    h = torch.cat([cls, t], 1) + sd['embeddings.position_embeddings']

    per_layer = []
    for i in range(hp['layers']):
        pre = f'encoder.layer.{i}.'
        # Transformer.forward — vit/vit.py:133-149 (LayerNorm eps 1e-12: :126,:130)
        y = layernorm(h, sd[pre + 'layernorm_before.weight'], sd[pre + 'layernorm_before.bias'], 1e-12)
        # MultiHeadAttention.forward — vit/vit.py:97-111, heads one at a time
        ctx = torch.empty_like(y)
        for hd in range(hp['heads']):
            hp_ = f'{pre}attention.attention.{hd}.'
            # SelfAttention.forward — vit/vit.py:56-74
            q = matmul(y, sd[hp_ + 'query.weight'], sd[hp_ + 'query.bias'])
            k = matmul(y, sd[hp_ + 'key.weight'], sd[hp_ + 'key.bias'])
            v = matmul(y, sd[hp_ + 'value.weight'], sd[hp_ + 'value.bias'])
            dh = q.shape[-1]
            scores = (q @ k.transpose(1, 2)) * (1 / math.sqrt(dh))   # matmul3.py:105-106
            ctx[:, :, hd * dh:(hd + 1) * dh] = softmax(scores) @ v
        attn = matmul(ctx, sd[pre + 'attention.output.weight'], sd[pre + 'attention.output.bias'])
        res = attn + h                                               # add.py:60-65
        y = layernorm(res, sd[pre + 'layernorm_after.weight'], sd[pre + 'layernorm_after.bias'], 1e-12)
        y = matmul(y, sd[pre + 'intermediate.weight'], sd[pre + 'intermediate.bias'], 'gelu')
        y = matmul(y, sd[pre + 'output.weight'], sd[pre + 'output.bias'])
        h = y + res
        per_layer.append(h)
    out = layernorm(h, sd['layernorm.weight'], sd['layernorm.bias'], 1e-12)   # vit/vit.py:245
    return (out, per_layer) if return_layers else out
