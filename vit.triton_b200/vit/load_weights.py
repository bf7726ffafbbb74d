"""HuggingFace ViT state-dict -> this package's module tree.

Same two entry points as the reference (vit/load_weights.py:11-37,39-62) with the same signatures
and the same resulting tensors for ViT-B/16, but the number of heads, the head width and the number
of layers are read from the destination state-dict instead of being hard-coded (64-wide heads:
load_weights.py:29,31; 12 layers: utils.py:53), so ViT-L/16 and ViT-H/14 load correctly too.
"""
import re
from typing import Dict

import torch

_HEAD_KEY = re.compile(r"^encoder\.layer\.(\d+)\.attention\.attention\.(\d+)\.(query|key|value)\.(weight|bias)$")


def map_attn_layers(source_layer_num: str, source_proj: str, source_type: str,
                    source_tensor: torch.Tensor, dest_state_dict: dict) -> Dict:
    """Scatter one HF query/key/value tensor of layer ``source_layer_num`` over the per-head entries.

    HF keeps all heads of a projection in one (D, D) ``nn.Linear`` weight (out, in) and a (D,) bias;
    here head h owns weight ``W.T[:, h*dh:(h+1)*dh]`` (in, out) and bias ``b[h*dh:(h+1)*dh]``.
    """
    transposed = source_tensor.T.contiguous() if source_type == 'weight' else None
    for name, dest in dest_state_dict.items():
        m = _HEAD_KEY.match(name)
        if m is None:
            continue
        layer_num, head, proj, kind = m.group(1), int(m.group(2)), m.group(3), m.group(4)
        if layer_num != str(source_layer_num) or proj != source_proj or kind != source_type:
            continue
        width = dest.shape[-1]
        lo, hi = head * width, (head + 1) * width
        piece = transposed[:, lo:hi] if kind == 'weight' else source_tensor[lo:hi]
        assert piece.shape == dest.shape, \
            f"Shape mismatch while mapping {source_proj}.{source_type} of layer {source_layer_num} to {name}: {piece.shape} vs {dest.shape}"
        dest_state_dict[name] = piece.clone()
    return dest_state_dict


def map_non_attn_layers(source_state_dict: dict, dest_state_dict: dict, weight_mapping: Dict) -> Dict:
    """Copy embeddings, LayerNorms and dense layers.  Dense weights ('output' / 'intermediate' /
    'pooler' / 'classifier' in the destination name) are transposed because HF stores (out, in) and this
    tree stores (in, out).  'pooler.*' only exists in the destination for ``VIT(add_pooling_layer=True)``; the
    reference maps the key (utils.py:63-64) and always skips it."""
    for src_name, tensor in source_state_dict.items():
        dst_name = weight_mapping.get(src_name)
        if not dst_name or dst_name not in dest_state_dict:
            continue
        dense = ('output' in dst_name) or ('intermediate' in dst_name) or ('pooler' in dst_name) \
            or ('classifier' in dst_name)
        dest_state_dict[dst_name] = tensor.t().clone() if (dense and tensor.dim() == 2) else tensor.clone()
    return dest_state_dict
