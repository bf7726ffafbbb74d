"""ViT encoder — host-side module tree with the reference's public API (vit/vit.py:25-247).

Class names, constructor arguments, parameter names / shapes (hence state-dict keys) and forward
signatures are those of cmeraki/vit.triton so this file drops in for the reference's; what runs
underneath is different.  Each ``Transformer`` block executes as SEVEN launches (six with the default
LayerNorm fold: ``layernorm_before`` of blocks 1.. runs inside the QKV GEMM's epilogue) of hand-written
sm_100a kernels over the flattened (B*N, D) activation:

    LN1 -> QKV GEMM (all heads, one launch) -> fused attention -> out-proj GEMM (+bias +residual)
        -> LN2 -> fc1 GEMM (+bias +GELU) -> fc2 GEMM (+bias +residual)

instead of the reference's 79 Triton launches + 24 torch copies per block.  The
per-head ``SelfAttention`` / ``LinearWithBias`` modules still exist, own the parameters and can be
run stand-alone (``set_fused(False)`` runs the whole model that way, which is what per-module
forward-hook comparisons against HuggingFace need).

Differences from the reference that are deliberate: the patch projection has
``hidden_dim`` output channels (the reference passes ``patch_dim``, equal only for ViT-B/16), the MLP
width can be set, and dtype / device follow the parameters instead of module-level globals.
"""
import math
import os
from typing import Optional

import torch
from torch import nn

from . import packing
from .kernels import (
    matmul,
    softmax,
    add,
    matmul3,
    LayerNormTriton,
    Conv2DTriton,
    flash_attention,
)
from .kernels import _lib
from .kernels.layernorm import layernorm

_FUSED = True
# Both LayerNorms of a block are folded into the GEMM that consumes them: the normalisation is applied
# per row in that GEMM's epilogue (zero-sum folded weights, packing.fold_layernorm / csrc/ln_fold.cu) and the row
# statistics come out of the epilogue of the GEMM that PRODUCES the row (fc2 of the previous block /
# the patch embedding for layernorm_before, the out-proj for layernorm_after).  Measured in one process
# under the sustained power cap (tools/fold_ab.py): 9.19 / 9.01 / 8.76 ms per forward with no fold / only
# layernorm_before / both — 24 launches and 3.7 GB of activation traffic less per forward.
_FOLD_LN = True
_FOLD_LN_MLP = True    # layernorm_after -> fc1 fold (statistics from the out-proj epilogue), see set_layernorm_folding


# Optional FP8 path (off by default and off the bf16 headline metric; VT_FP8=1 or set_fp8(True)):
# the QKV, fc1 and fc2 GEMMs of every block run on e4m3 operands (tcgen05.mma kind::f8f6f4): the LayerNorms write
# e4m3 directly, fc1's epilogue writes e4m3 for fc2, weights are quantised per output channel at pack time; the
# attention, the out-projection, the residual stream and the final LayerNorm stay bf16.
_FP8 = os.environ.get("VT_FP8", "0") == "1"


def set_fp8(enabled: bool) -> None:
    global _FP8
    _FP8 = bool(enabled)


def fp8_enabled() -> bool:
    return _FP8


def set_fused(enabled: bool) -> None:
    """Toggle the fused per-block execution (default on).  Off = one kernel entry point per
    reference op, heads processed one by one, every sub-module's forward() is really called."""
    global _FUSED
    _FUSED = bool(enabled)


def fused_enabled() -> bool:
    return _FUSED


def set_layernorm_folding(enabled: bool, mlp: bool = True) -> None:
    """Toggle folding of the LayerNorms into the GEMM epilogues (default: both ON, see the note at _FOLD_LN;
    only affects the fused bf16 path).  ``mlp=False`` keeps layernorm_after as a kernel (6 launches per
    block), ``enabled=False`` both (7 launches per block instead of 5)."""
    global _FOLD_LN, _FOLD_LN_MLP
    _FOLD_LN = bool(enabled)
    _FOLD_LN_MLP = bool(enabled and mlp)


class LinearWithBias(nn.Module):
    """y = act(x @ weight + bias); weight is stored (in, out) like the reference (vit.py:25-35)."""

    def __init__(self, input_dim: int, output_dim: int, activation: Optional[str] = None):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(input_dim, output_dim))
        self.bias = nn.Parameter(torch.zeros(output_dim))
        self.activation = activation

    def forward(self, x) -> torch.Tensor:
        return matmul(x, self.weight, self.bias, self.activation)


class SelfAttention(nn.Module):
    """One attention head (reference vit.py:38-74).  Stand-alone forward = three projections,
    scaled scores, softmax, weighted sum — each through its kernel entry point."""

    def __init__(self, d_in: int, d_out: int, dropout: int = 0):
        super().__init__()
        self.d_in = d_in
        self.d_out = d_out
        self.query = LinearWithBias(self.d_in, self.d_out)
        self.key = LinearWithBias(self.d_in, self.d_out)
        self.value = LinearWithBias(self.d_in, self.d_out)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        q = self.query(x)
        k_t = self.key(x).transpose(1, 2).contiguous()
        v = self.value(x)
        probs = softmax(matmul3(q, k_t, apply_scaling=True, scale_factor=1 / math.sqrt(self.d_out)))
        return matmul3(probs, v)


class MultiHeadAttention(packing.PackedMixin, nn.Module):
    """All heads + output projection (reference vit.py:77-111)."""

    def __init__(self, num_heads: int, d_in: int, d_out: int):
        super().__init__()
        assert d_in % num_heads == 0, f'Input dimension should be equally divided amongst all heads. d_in%num_heads needs to be 0. Current: {d_in%num_heads}'
        assert d_in / num_heads == d_out, f'`d_out` is not equal to `d_in/num_heads`. Current: {d_in/num_heads}, {d_out}'
        self.num_heads = num_heads
        self.d_in = d_in
        self.d_out = d_out
        self.attention = nn.ModuleList([SelfAttention(d_in=d_in, d_out=d_out) for _ in range(self.num_heads)])
        self.output = LinearWithBias(self.d_in, self.d_in)

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: (B, N, D).  ``residual`` (fused path only) is added in the out-projection epilogue."""
        if not (_FUSED and x.is_cuda):
            assert residual is None
            heads_out = torch.empty_like(x)
            for i, head in enumerate(self.attention):
                heads_out[:, :, i * self.d_out:(i + 1) * self.d_out] = head(x)
            return self.output(heads_out)

        pk = self.packed()
        B, N, D = x.shape
        x = x.contiguous()
        qkv = packing.linear(x, pk.wqkv, pk.bqkv)
        ctx = flash_attention(qkv, self.num_heads, 1.0 / math.sqrt(self.d_out))
        return packing.linear(ctx, pk.wo, pk.bo, residual=residual)

    def _build_packed(self):
        return packing.pack_attention(self)


class Transformer(packing.PackedMixin, nn.Module):
    """Pre-LN encoder block (reference vit.py:114-149): res = x + MHA(LN1(x));
    out = res + W2 . GELU(W1 . LN2(res)).  LayerNorm eps is 1e-12 like HF ViT."""

    def __init__(self, num_heads: int, d_in: int, d_out: int, mlp_dim: Optional[int] = None):
        super().__init__()
        self.num_heads = num_heads
        self.d_in = d_in
        self.d_out = d_out
        self.mlp_dim = 4 * self.d_in if mlp_dim is None else mlp_dim

        self.layernorm_before = LayerNormTriton(self.d_in, eps=1e-12)
        self.attention = MultiHeadAttention(self.num_heads, self.d_in, self.d_out)
        self.intermediate = LinearWithBias(self.d_in, self.mlp_dim, activation='gelu')
        self.output = LinearWithBias(self.mlp_dim, self.d_in)
        self.layernorm_after = LayerNormTriton(self.d_in, eps=1e-12)

    def forward(self, x):
        if not (_FUSED and x.is_cuda):
            res = add(self.attention(self.layernorm_before(x)), x)
            out = self.output(self.intermediate(self.layernorm_after(res)))
            return add(out, res)

        pk = self.packed()
        x = x.contiguous()
        res = self.attention(self.layernorm_before(x), residual=x)
        mid = packing.linear(self.layernorm_after(res), pk.w1, pk.b1, gelu=True)
        return packing.linear(mid, pk.w2, pk.b2, residual=res)

    def _packed_sources(self):
        return [self.intermediate.weight, self.intermediate.bias, self.output.weight, self.output.bias]

    def _build_packed(self):
        return packing.pack_mlp(self)

    def _packed_sources_folded(self):
        return list(self.parameters())

    def _build_packed_folded(self):
        return packing.pack_block_folded(self)

    def _packed_sources_fp8(self):
        return list(self.parameters())

    def _build_packed_fp8(self):
        return packing.pack_block_fp8(self)

    def forward_fp8(self, x: torch.Tensor) -> torch.Tensor:
        """Block forward with the QKV / fc1 / fc2 GEMMs on e4m3 operands (7 launches; see the note at _FP8)."""
        att = self.attention.packed()
        mlp = self.packed()
        q8 = self.packed("fp8")
        x = x.contiguous()
        qkv = packing.linear_fp8(packing.layernorm_fp8(x, self.layernorm_before), q8.wqkv8, q8.cqkv, att.bqkv)
        ctx = flash_attention(qkv, self.num_heads, 1.0 / math.sqrt(self.d_out))
        res = packing.linear(ctx, att.wo, att.bo, residual=x)
        mid8 = packing.linear_fp8(packing.layernorm_fp8(res, self.layernorm_after), q8.w18, q8.c1, mlp.b1, gelu=True,
                                  out_fp8=True, out_scale=packing.FP8_MID_SCALE)
        return packing.linear_fp8(mid8, q8.w28, q8.c2, mlp.b2, residual=res)

    def forward_folded(self, x: torch.Tensor, ln1_stats: Optional[torch.Tensor]):
        """Block forward with layernorm_before folded into the QKV GEMM (bf16 only).

        ``ln1_stats``: (M, D/128, 2) fp32 per-128-column (sum, M2) partials of x's rows, written by the
        previous block's last GEMM (None for the first block: layernorm_before then runs as a
        kernel).  Returns (output, statistics of the output rows).  5 launches per block (layernorm_after folded into
        fc1 as well, the default) instead of 7."""
        att = self.attention.packed()
        mlp = self.packed()
        x = x.contiguous()
        if ln1_stats is None:
            qkv = packing.linear(self.layernorm_before(x), att.wqkv, att.bqkv)
        else:
            pk = self.packed("folded")
            qkv = packing.linear_ln(x, pk.wqkv, pk.bqkv, pk.cqkv, ln1_stats, self.layernorm_before.eps)
        ctx = flash_attention(qkv, self.num_heads, 1.0 / math.sqrt(self.d_out))
        if _FOLD_LN_MLP:
            pk = self.packed("folded")
            stats2 = torch.empty((x.shape[0] * x.shape[1], self.d_in // packing.STATS_COLS, 2), device=x.device,
                                 dtype=torch.float32)
            res = packing.linear_res_stats(ctx, att.wo, att.bo, x, stats2)
            mid = packing.linear_ln(res, pk.w1, pk.b1, pk.c1, stats2, self.layernorm_after.eps, gelu=True)
        else:
            res = packing.linear(ctx, att.wo, att.bo, residual=x)
            mid = packing.linear(self.layernorm_after(res), mlp.w1, mlp.b1, gelu=True)
        stats = torch.empty((x.shape[0] * x.shape[1], self.d_in // packing.STATS_COLS, 2), device=x.device, dtype=torch.float32)
        out = packing.linear_res_stats(mid, mlp.w2, mlp.b2, res, stats)
        return out, stats


class Encoder(nn.Module):
    """Stack of blocks (reference vit.py:152-170)."""

    def __init__(self, num_layers: int, num_heads: int, hidden_dim: int, d_out: int,
                 mlp_dim: Optional[int] = None):
        super().__init__()
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.hidden_dim = hidden_dim
        self.d_out = d_out
        self.layer = nn.ModuleList(
            Transformer(num_heads=self.num_heads, d_in=self.hidden_dim, d_out=d_out, mlp_dim=mlp_dim)
            for _ in range(self.num_layers)
        )

    def fp8_active(self, x) -> bool:
        return bool(_FUSED and _FP8 and len(self.layer) > 0 and x.numel() > 0 and
                    packing.fp8_supported(x, self.hidden_dim, self.layer[0].mlp_dim))

    def folding_active(self, x) -> bool:
        return bool(_FUSED and _FOLD_LN and not self.fp8_active(x) and len(self.layer) > 0 and x.numel() > 0 and
                    packing.folding_supported(x, self.hidden_dim, self.layer[0].mlp_dim))

    def forward(self, x, ln1_stats: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``ln1_stats``: row statistics of x written by the patch-embedding epilogue (see
        ``Embeddings.forward``); without them block 0's layernorm_before runs as a kernel."""
        if self.fp8_active(x):
            for layer in self.layer:
                x = layer.forward_fp8(x)
            return x
        if self.folding_active(x):
            # layernorm_before folded into the QKV GEMM; row statistics of each block's output are
            # produced by its last GEMM's epilogue
            for layer in self.layer:
                x, ln1_stats = layer.forward_folded(x, ln1_stats)
            return x
        for layer in self.layer:
            x = layer(x)
        return x


class Embeddings(packing.PackedMixin, nn.Module):
    """Patch projection + CLS token + position embeddings (reference vit.py:173-200)."""

    def __init__(self, patch_size, num_patches, patch_dim, hidden_dim, channels: int = 3):
        super().__init__()
        self.patch_size = patch_size
        self.num_patches = num_patches
        self.patch_dim = patch_dim
        self.hidden_dim = hidden_dim
        self.channels = channels

        self.cls_token = nn.Parameter(torch.zeros((1, 1, hidden_dim)))
        self.position_embeddings = nn.Parameter(torch.zeros(1, num_patches + 1, hidden_dim))
        self.projection = Conv2DTriton(
            in_channels=channels,
            out_channels=hidden_dim,
            kernel_size=(self.patch_size, self.patch_size)
        )

    def forward(self, x, stats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``stats_out`` (optional, fused bf16 path only): (B * N, D/128, 2) fp32 buffer that receives the
        row statistics of the output for the LayerNorm folded into block 0's QKV GEMM."""
        if not (_FUSED and x.is_cuda):
            assert stats_out is None
            tokens = self.projection(x.to(self.projection.weight.dtype)).flatten(2).transpose(1, 2)
            out = torch.empty((x.shape[0], self.num_patches + 1, self.hidden_dim), device=x.device,
                              dtype=tokens.dtype)
            out[:, 1:, :] = tokens
            _lib.call("vt_embed_finalize", out.data_ptr(), self.position_embeddings.data_ptr(),
                      self.cls_token.data_ptr(), out.shape[0], out.shape[1], out.shape[2],
                      _lib.dtype_code(out), _lib.stream_ptr(out))
            return out
        return packing.patch_embed(self, x, stats_out)

    def forward_uint8(self, x, image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5),
                      rescale_factor: float = 1.0 / 255.0, stats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Raw uint8 NHWC pixels (B, H, W, C) -> embeddings: the image processor's rescale and
        normalisation (defaults: HF ``ViTImageProcessor`` of google/vit-base-patch16-224) are folded
        into the patch projection, see ``packing.pack_embeddings_u8``."""
        return packing.patch_embed_u8(self, x, image_mean, image_std, rescale_factor, stats_out)

    def _packed_sources(self):
        return [self.cls_token, self.position_embeddings, self.projection.weight, self.projection.bias]

    def _build_packed(self):
        return packing.pack_embeddings(self)


def _head_on_cls(hidden: torch.Tensor, dense: "LinearWithBias", act: int) -> torch.Tensor:
    """(B, n_out) = act(dense(hidden[:, 0, :])): ONE tensor-core launch (``vt_bgemm``) that reads the CLS
    rows in place through the row stride of its tensor map (no pooling kernel, no copy) and applies bias
    and activation in the epilogue; fp32 models pack the rows into split bf16 pieces first."""
    from .kernels import bgemm as bg
    from .kernels.matmul import _cached
    assert hidden.is_cuda and hidden.dim() == 3 and hidden.stride(2) == 1
    B, _, D = hidden.shape
    w = dense.weight                                       # (in, out)
    n_out = w.shape[1]
    out = torch.empty((B, n_out), device=hidden.device, dtype=hidden.dtype)
    if B == 0:
        return out
    pieces = 1 if hidden.dtype == torch.bfloat16 else bg.split_pieces()
    wp = _cached(w, f"packed{pieces}", lambda p_: bg.pack(p_, p_.data_ptr(), n_out, D, 1, 1,
                                                          (0, 0, p_.stride(1), p_.stride(0)), pieces, pattern=1)[0])
    bias32 = _cached(dense.bias, "f32", lambda b: b.detach().float().contiguous())
    bg.dense_rows(hidden, hidden.data_ptr(), B, D, hidden.stride(0), wp, pieces, n_out, bias32, act, out,
                  out.data_ptr(), n_out)
    return out


class Pooler(nn.Module):
    """HF ``ViTPooler`` (modeling_vit.py:461-474): tanh(dense(CLS hidden state)).  The reference's
    loader already maps ``pooler.dense.{weight,bias}`` (vit/utils.py:63-64) but the reference model has
    no such module; ``VIT(..., add_pooling_layer=True)`` adds it.  On the GPU the dense layer and the tanh
    are one tcgen05 GEMM launch with a tanh epilogue over the CLS rows (``_head_on_cls``); weight is (in, out)."""

    def __init__(self, hidden_dim: int):
        super().__init__()
        self.dense = LinearWithBias(hidden_dim, hidden_dim)

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        if hidden_states.is_cuda:
            from .kernels import bgemm as bg
            return _head_on_cls(hidden_states, self.dense, bg.ACT_TANH)
        return torch.tanh(self.dense(hidden_states[:, :1, :].contiguous())[:, 0, :])


class VIT(nn.Module):
    """ViT encoder returning the final-LayerNorm hidden states (B, N, D) — no pooler, like the
    reference (vit.py:203-247) and HF ``ViTModel(add_pooling_layer=False).last_hidden_state``."""

    def __init__(
        self,
        height: int,
        width: int,
        channels: int,
        patch_size: int,
        hidden_dim: int,
        num_heads: int,
        num_layers: int,
        mlp_dim: Optional[int] = None,
        add_pooling_layer: bool = False,
        num_labels: Optional[int] = None,
    ):
        super().__init__()
        assert height == width, "Height and width should be the same"
        assert height % patch_size == 0, "Height should be divisible by the patch size"
        assert width % patch_size == 0, "Width should be divisible by the patch size"

        self.height = height
        self.width = width
        self.channels = channels
        self.patch_size = patch_size
        self.hidden_dim = hidden_dim
        self.num_heads = num_heads
        self.num_layers = num_layers

        assert self.hidden_dim % self.num_heads == 0, f"Hidden dimension should be divisible by number of heads, provided: {self.hidden_dim} {self.num_heads}"

        num_patches = (self.height // self.patch_size) * (self.width // self.patch_size)
        patch_dim = self.patch_size * self.patch_size * self.channels
        d_out = self.hidden_dim // self.num_heads

        self.embeddings = Embeddings(patch_size=self.patch_size, num_patches=num_patches, patch_dim=patch_dim,
                                     hidden_dim=self.hidden_dim, channels=self.channels)
        self.encoder = Encoder(num_layers=self.num_layers, num_heads=self.num_heads, hidden_dim=self.hidden_dim,
                               d_out=d_out, mlp_dim=mlp_dim)
        self.layernorm = LayerNormTriton(dim=self.hidden_dim, eps=1e-12)
        # optional HF pooler (tanh(dense(CLS))): state-dict keys pooler.dense.{weight,bias}
        self.pooler = Pooler(self.hidden_dim) if add_pooling_layer else None
        # optional HF ViTForImageClassification head (Linear on the CLS row of the final hidden states):
        # state-dict keys classifier.{weight,bias}
        self.classifier = LinearWithBias(self.hidden_dim, num_labels) if num_labels else None

    def invalidate_packed(self) -> None:
        """Drop every derived weight (packed / LayerNorm-folded / uint8-scaled operands of all sub-modules
        and the K-major copies the ``matmul`` entry point keeps).  Needed only after writes that go THROUGH
        ``param.data`` (``p.data.copy_()``, HF-style re-initialisation, EMA): those change neither the
        storage pointer nor the version counter the caches are keyed on; every other kind of update is
        noticed automatically."""
        from .kernels.matmul import clear_derived_cache
        for m in self.modules():
            if isinstance(m, packing.PackedMixin):
                m.invalidate_packed()
            m.__dict__.pop("_u8_pack", None)
        clear_derived_cache()

    @property
    def device(self) -> torch.device:
        return self.layernorm.weight.device

    @property
    def dtype(self) -> torch.dtype:
        return self.layernorm.weight.dtype

    def forward(self, x):
        assert x.shape[1:] == (self.channels, self.height, self.width), f"Image size {x.shape[1:]} not matching with the model input size: {self.channels, self.height, self.width}"
        stats = self._embed_stats(x)
        x = self.embeddings(x, stats) if stats is not None else self.embeddings(x)
        x = self.encoder(x, stats) if stats is not None else self.encoder(x)
        x = self.layernorm(x)
        return x

    def _embed_stats(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        """Buffer for the row statistics of the embeddings when block 0's layernorm_before is folded
        into its QKV GEMM (fused bf16 tensor-core path, hidden size a multiple of 128); None otherwise."""
        w = self.embeddings.projection.weight
        batch = x.shape[0]
        if not (x.is_cuda and w.is_cuda and w.dtype == torch.bfloat16 and batch > 0 and self.hidden_dim % 8 == 0
                and x.dtype in (torch.float32, torch.bfloat16, torch.uint8)):
            return None
        probe = torch.empty((1,), device=x.device, dtype=w.dtype)
        if not self.encoder.folding_active(probe):
            return None
        n_tok = self.embeddings.num_patches + 1
        return torch.empty((batch * n_tok, self.hidden_dim // packing.STATS_COLS, 2), device=x.device,
                           dtype=torch.float32)

    def forward_uint8(self, x, image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5),
                      rescale_factor: float = 1.0 / 255.0):
        """Forward from RAW uint8 NHWC pixels (B, H, W, C) — the step in front of the reference's
        ``forward``: rescale + normalise (HF ``ViTImageProcessor``) run inside the patch-embedding
        kernel, so the host -> device copy is one byte per pixel value."""
        assert tuple(x.shape[1:]) == (self.height, self.width, self.channels), f"Image size {x.shape[1:]} not matching with the model input size: {self.height, self.width, self.channels}"
        stats = self._embed_stats(x)
        x = self.embeddings.forward_uint8(x, image_mean, image_std, rescale_factor, stats)
        x = self.encoder(x, stats) if stats is not None else self.encoder(x)
        x = self.layernorm(x)
        return x

    def pooler_output(self, x) -> torch.Tensor:
        """HF ``ViTModel(...).pooler_output``: tanh(dense(final hidden state of CLS)), (B, D)."""
        assert self.pooler is not None, "Model was built without add_pooling_layer=True"
        hidden = self.forward_uint8(x) if x.dtype == torch.uint8 else self.forward(x)
        return self.pooler(hidden)

    def logits(self, x) -> torch.Tensor:
        """Class logits (B, num_labels) = classifier(final hidden states[:, 0]) — HF
        ``ViTForImageClassification(...).logits``; needs ``VIT(..., num_labels=...)``.  One tensor-core
        launch over the CLS rows of the final hidden states."""
        assert self.classifier is not None, "VIT was built without a classifier head (num_labels)"
        hidden = self.forward_uint8(x) if x.dtype == torch.uint8 else self.forward(x)
        from .kernels import bgemm as bg
        return _head_on_cls(hidden, self.classifier, bg.ACT_NONE)

    def pooled(self, x) -> torch.Tensor:
        """CLS row of the final hidden states, (B, D): the tensor the data-parallel wrapper gathers.
        uint8 (B, H, W, C) inputs take the fused raw-pixel path."""
        hidden = self.forward_uint8(x) if x.dtype == torch.uint8 else self.forward(x)
        out = torch.empty((hidden.shape[0], hidden.shape[2]), device=hidden.device, dtype=hidden.dtype)
        _lib.call("vt_pool_cls", hidden.data_ptr(), out.data_ptr(), hidden.shape[0], hidden.shape[2],
                  hidden.stride(0), _lib.dtype_code(hidden), _lib.stream_ptr(hidden))
        return out
