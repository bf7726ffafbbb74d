"""Weight packing for the fused execution path.

The module tree keeps the reference's parameter layout (per-head (D, dh) query/key/value matrices,
(in, out) dense weights — vit/vit.py:25-35,38-53; 990 state-dict keys for ViT-B).  The tensor-core
kernels want something else: ONE K-major [3D, D] matrix for Q, K and V of all heads, K-major
[out, in] matrices for the other dense layers, fp32 biases, and the position embedding pre-added to
the CLS token / conv bias.  This module derives those once and caches them per module; the cache is
dropped when the module is moved / cast (``_apply``), when a state-dict is loaded, or when any
source parameter's storage pointer or version counter changes (in-place ops, optimizer steps,
``p.data = new``).  Writes THROUGH ``p.data`` (``p.data.copy_()``, ``p.data.normal_()``, EMA updates)
bump no version counter and cannot be seen without reading the device: call ``VIT.invalidate_packed()``
after them.
"""
from types import SimpleNamespace
from typing import Optional

import torch

import os

from .kernels import _lib, bgemm as bg


class PackedMixin:
    """Gives a module lazily built, automatically invalidated ``packed(kind)`` namespaces.

    ``kind`` selects a builder ``_build_packed[_<kind>]`` and its source list
    ``_packed_sources[_<kind>]`` (default: every parameter of the module)."""

    def _packed_sources(self):
        return list(self.parameters())

    @staticmethod
    def _packed_key(params):
        return tuple((p.data_ptr(), p._version) for p in params)

    def packed(self, kind: str = ""):
        states = self.__dict__.setdefault("_packed_states", {})
        state = states.get(kind)
        if state is not None:
            params, key, value = state
            if self._packed_key(params) == key:
                return value
        suffix = f"_{kind}" if kind else ""
        params = getattr(self, "_packed_sources" + suffix, self._packed_sources)()
        with torch.no_grad():
            value = getattr(self, "_build_packed" + suffix)()
        states[kind] = (params, self._packed_key(params), value)
        return value

    def invalidate_packed(self):
        self.__dict__.pop("_packed_states", None)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_packed()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate_packed()
        return super()._load_from_state_dict(*args, **kwargs)


def _nk(weight_in_out: torch.Tensor) -> torch.Tensor:
    """(in, out) -> contiguous [out, in]: the K-major operand layout."""
    return weight_in_out.detach().t().contiguous()


def pack_attention(mha) -> SimpleNamespace:
    heads = mha.attention
    wq = torch.cat([h.query.weight.detach() for h in heads], dim=1)
    wk = torch.cat([h.key.weight.detach() for h in heads], dim=1)
    wv = torch.cat([h.value.weight.detach() for h in heads], dim=1)
    bq = torch.cat([h.query.bias.detach() for h in heads])
    bk = torch.cat([h.key.bias.detach() for h in heads])
    bv = torch.cat([h.value.bias.detach() for h in heads])
    return SimpleNamespace(
        wqkv=_nk(torch.cat([wq, wk, wv], dim=1)),          # [3D, D]
        bqkv=torch.cat([bq, bk, bv]).float().contiguous(),  # [3D] fp32
        wo=_nk(mha.output.weight),                          # [D, D]
        bo=mha.output.bias.detach().float().contiguous(),
    )


def pack_mlp(block) -> SimpleNamespace:
    return SimpleNamespace(
        w1=_nk(block.intermediate.weight),                  # [F, D]
        b1=block.intermediate.bias.detach().float().contiguous(),
        w2=_nk(block.output.weight),                        # [D, F]
        b2=block.output.bias.detach().float().contiguous(),
    )


def fold_layernorm(w_nk: torch.Tensor, bias32: torch.Tensor, ln, zero_sum: bool = True) -> tuple:
    """Fold y = LN(x) into the dense layer that consumes it (one launch of ``vt_ln_fold``, csrc/ln_fold.cu):

        LN(x) @ W^T + b = rstd * (x @ (W * gamma)^T) - rstd * mean * colsum(W * gamma) + (b + W @ beta).

    ``zero_sum`` (the default packing) moves the mean term INTO the weights: every row of W * gamma is
    shifted by its own mean, so  x @ W'^T = x @ (W * gamma)^T - mean(x) * colsum  comes out of the GEMM
    itself and the epilogue is  rstd * acc + (b + W @ beta)  — no column-sum operand, one FMA per element
    less; the bf16 rounding residue of each row sum (~3e-3 for a 768-wide row) is cancelled in the kernel
    by re-rounding the elements closest to a rounding tie the other way (left: ~1e-6).
    Returns (W' bf16 [N, K], b' fp32 [N], colsum fp32 [N] or None for the zero-sum form)."""
    assert w_nk.is_cuda and w_nk.dtype == torch.bfloat16 and w_nk.is_contiguous(), \
        "LayerNorm folding runs on CUDA bf16 K-major weights (there is no CPU path)"
    N, K = w_nk.shape
    gamma = ln.weight.detach().float().contiguous()
    beta = ln.bias.detach().float().contiguous()
    bias32 = bias32.contiguous()
    w_out = torch.empty_like(w_nk)
    b_out = torch.empty((N,), device=w_nk.device, dtype=torch.float32)
    colsum = None if zero_sum else torch.empty((N,), device=w_nk.device, dtype=torch.float32)
    _lib.call("vt_ln_fold", w_nk.data_ptr(), K, bias32.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
              w_out.data_ptr(), K, b_out.data_ptr(), _lib.ptr(colsum), N, K, 1 if zero_sum else 0,
              _lib.stream_ptr(w_nk))
    return w_out, b_out, colsum


def pack_block_folded(block) -> SimpleNamespace:
    """Both LayerNorms of one encoder block folded into the GEMMs that consume them, zero-sum form:
    layernorm_before -> fused QKV weights, layernorm_after -> fc1 weights (two ``vt_ln_fold`` launches)."""
    att = block.attention.packed()
    mlp = block.packed()
    wqkv, bqkv, _ = fold_layernorm(att.wqkv, att.bqkv, block.layernorm_before)
    w1, b1, _ = fold_layernorm(mlp.w1, mlp.b1, block.layernorm_after)
    return SimpleNamespace(wqkv=wqkv, bqkv=bqkv, cqkv=None, w1=w1, b1=b1, c1=None)


# ------------------------------------------------------------------------------------------------ FP8 (optional)
FP8_ACT_SCALE = 16.0     # LayerNorm outputs are quantised as e4m3(y * 16): |LN(x)| <= sqrt(dim - 1), 448 / 16 = 28
FP8_MID_SCALE = 8.0      # GELU outputs (fc2's operand) as e4m3(g * 8), saturating at |g| = 56


def fp8_supported(x: torch.Tensor, dim: int, mlp_dim: int) -> bool:
    return x.is_cuda and x.dtype == torch.bfloat16 and dim % 16 == 0 and mlp_dim % 128 == 0


def quantize_weight_fp8(w_nk: torch.Tensor) -> tuple:
    """bf16 [N, K] K-major weight -> (e4m3 bytes [N, K], fp32 per-output-channel scales [N])."""
    assert w_nk.is_cuda and w_nk.dtype == torch.bfloat16 and w_nk.is_contiguous()
    N, K = w_nk.shape
    w8 = torch.empty((N, K), device=w_nk.device, dtype=torch.uint8)
    scales = torch.empty((N,), device=w_nk.device, dtype=torch.float32)
    _lib.call("vt_quantize_rows_fp8", w_nk.data_ptr(), K, w8.data_ptr(), K, scales.data_ptr(), N, K,
              _lib.stream_ptr(w_nk))
    return w8, scales


def pack_block_fp8(block) -> SimpleNamespace:
    """e4m3 weights of the QKV / fc1 / fc2 layers of one block with their per-channel dequantisation scales,
    the activation scales already divided out (the GEMM epilogue is acc * colscale + bias)."""
    att = block.attention.packed()
    mlp = block.packed()
    wqkv8, sqkv = quantize_weight_fp8(att.wqkv)
    w18, s1 = quantize_weight_fp8(mlp.w1)
    w28, s2 = quantize_weight_fp8(mlp.w2)
    return SimpleNamespace(wqkv8=wqkv8, cqkv=(sqkv / FP8_ACT_SCALE).contiguous(), w18=w18,
                           c1=(s1 / FP8_ACT_SCALE).contiguous(), w28=w28, c2=(s2 / FP8_MID_SCALE).contiguous())


def layernorm_fp8(x: torch.Tensor, ln, scale: float = FP8_ACT_SCALE) -> torch.Tensor:
    """(B, N, D) bf16 -> e4m3(LN(x) * scale) as uint8 (B, N, D): LayerNorm with the GEMM operand's quantisation fused."""
    B, N, D = x.shape
    out = torch.empty((B, N, D), device=x.device, dtype=torch.uint8)
    if out.numel():
        _lib.call("vt_layernorm_fp8", x.data_ptr(), ln.weight.data_ptr(), ln.bias.data_ptr(), out.data_ptr(), B * N, D,
                  D, D, float(ln.eps), float(scale), _lib.stream_ptr(x))
    return out


def linear_fp8(x8: torch.Tensor, w8: torch.Tensor, colscale: torch.Tensor, bias32: torch.Tensor, gelu: bool = False,
               residual: Optional[torch.Tensor] = None, out_fp8: bool = False, out_scale: float = 1.0) -> torch.Tensor:
    """out = act((x8 @ w8^T) * colscale + bias) (+ residual) on tcgen05.mma kind::f8f6f4; x8 / w8 are e4m3 bytes."""
    B, N, K = x8.shape
    n_out = w8.shape[0]
    out = torch.empty((B, N, n_out), device=x8.device, dtype=torch.uint8 if out_fp8 else torch.bfloat16)
    if out.numel():
        _lib.call("vt_gemm_fp8", x8.data_ptr(), K, w8.data_ptr(), K, out.data_ptr(), n_out,
                  _lib.VT_E4M3 if out_fp8 else _lib.VT_BF16, bias32.data_ptr(), colscale.data_ptr(),
                  _lib.ptr(residual), n_out, B * N, n_out, K, 1 if gelu else 0, float(out_scale), _lib.stream_ptr(x8))
    return out


def _posb16(pos32: torch.Tensor, cls_token: torch.Tensor, bias32: torch.Tensor) -> torch.Tensor:
    """bf16 [N, D] position table of the two-launch patch embedding (``vt_patch_embed_gemm``): the GEMM adds the
    conv bias to EVERY row, so row 0 (CLS: a zero patch row) carries cls + pos[0] - bias."""
    t = pos32.clone()
    t[0] += cls_token.detach().float().reshape(-1) - bias32
    return t.to(torch.bfloat16).contiguous()


def pack_embeddings(emb) -> SimpleNamespace:
    w = emb.projection.weight.detach()
    D = w.shape[0]
    K = w[0].numel()
    kpad = (K + 7) // 8 * 8
    w2d = torch.zeros((D, kpad), device=w.device, dtype=w.dtype)
    w2d[:, :K] = w.reshape(D, K)
    pos = emb.position_embeddings.detach().float()[0]          # [N, D]
    posb = pos.clone()
    posb[0] += emb.cls_token.detach().float().reshape(-1)
    posb[1:] += emb.projection.bias.detach().float()
    bias32 = emb.projection.bias.detach().float().contiguous()
    return SimpleNamespace(w=w2d, ldw=kpad, K=K, posb=posb.contiguous(), bias32=bias32,
                           posb16=_posb16(pos, emb.cls_token, bias32))


def pack_embeddings_u8(emb, image_mean, image_std, rescale_factor: float) -> SimpleNamespace:
    """Patch projection for RAW uint8 NHWC pixels: the image processor's
    ``(x * rescale_factor - mean_c) / std_c`` (HF ViTImageProcessor: 1/255, 0.5, 0.5) is folded into
    the operands, so the kernel multiplies bytes:

        sum_k w[d,k] * (x_k * r - mean_c) / std_c  =  sum_k (w[d,k] * r / std_c) * x_k  -  sum_k w[d,k] * mean_c / std_c

    The weight is re-ordered from the reference's (c, i, j) to (i, j, c) — a patch row of an NHWC
    image is P*C contiguous bytes — and scaled in fp32 before the single rounding to bf16; the
    constant term joins the conv bias in the fp32 position/bias table."""
    w = emb.projection.weight.detach().float()                 # [D, C, P, P]
    D, C = w.shape[0], w.shape[1]
    mean = torch.as_tensor(image_mean, dtype=torch.float32, device=w.device).reshape(1, C, 1, 1)
    std = torch.as_tensor(image_std, dtype=torch.float32, device=w.device).reshape(1, C, 1, 1)
    w_scaled = w * (float(rescale_factor) / std)
    offset = (w * (mean / std)).sum(dim=(1, 2, 3))              # [D]
    K = w[0].numel()
    kpad = (K + 7) // 8 * 8
    w2d = torch.zeros((D, kpad), device=w.device, dtype=emb.projection.weight.dtype)
    w2d[:, :K] = w_scaled.permute(0, 2, 3, 1).reshape(D, K).to(w2d.dtype)
    pos = emb.position_embeddings.detach().float()[0]
    posb = pos.clone()
    posb[0] += emb.cls_token.detach().float().reshape(-1)
    posb[1:] += emb.projection.bias.detach().float() - offset
    bias32 = (emb.projection.bias.detach().float() - offset).contiguous()
    return SimpleNamespace(w=w2d, ldw=kpad, K=K, posb=posb.contiguous(), bias32=bias32,
                           posb16=_posb16(pos, emb.cls_token, bias32))


def patch_embed_u8(emb, x: torch.Tensor, image_mean, image_std, rescale_factor: float,
                   stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Raw uint8 NHWC pixels (B, S, S, C) -> token embeddings (B, n+1, D), rescale + normalise + patch
    projection + CLS + position embeddings in ONE kernel (bf16 models on the tensor-core path)."""
    w = emb.projection.weight
    assert x.is_cuda and x.dtype == torch.uint8 and x.dim() == 4, \
        f"Raw pixels need to be a CUDA uint8 (B, H, W, C) tensor, provided: {x.dtype}, {tuple(x.shape)}"
    assert w.dtype == torch.bfloat16 and emb.hidden_dim % 8 == 0, \
        "The fused uint8 input path runs on the bf16 tensor-core kernels only"
    B, S, S2, C = x.shape
    assert S == S2 and C == emb.channels, f"Image size {tuple(x.shape[1:])} not matching with the model input size"
    key = (tuple(float(m) for m in image_mean), tuple(float(v) for v in image_std), float(rescale_factor))
    cache = emb.__dict__.setdefault("_u8_pack", {})
    params = emb._packed_sources()
    stamp = emb._packed_key(params)
    hit = cache.get(key)
    if hit is None or hit[0] != stamp:
        with torch.no_grad():
            hit = (stamp, pack_embeddings_u8(emb, image_mean, image_std, rescale_factor))
        cache.clear()
        cache[key] = hit
    pk = hit[1]
    x = x.contiguous()
    n_tok = emb.num_patches + 1
    out = torch.empty((B, n_tok, emb.hidden_dim), device=x.device, dtype=w.dtype)
    if B == 0:
        return out
    stream = _lib.stream_ptr(x)
    step = 65535 * 128 // emb.num_patches
    for b0 in range(0, B, step):
        nb = min(step, B - b0)
        _embed_call(stats, b0, n_tok, emb.hidden_dim, pk, x[b0:].data_ptr(), _lib.VT_U8, out[b0:].data_ptr(), nb, C, S,
                    emb.patch_size, stream, x.device)
    return out


STATS_COLS = 128   # columns per (sum, M2) partial of the row statistics (VT_LN_STATS_COLS in include/vitb200.h)


def folding_supported(x: torch.Tensor, dim: int, mlp_dim: int) -> bool:
    """LayerNorm folding runs on the bf16 tensor-core GEMM only."""
    return x.is_cuda and x.dtype == torch.bfloat16 and dim % STATS_COLS == 0 and mlp_dim % 8 == 0


def linear_ln(x: torch.Tensor, w_fold: torch.Tensor, b_fold: torch.Tensor, colsum: Optional[torch.Tensor],
              rowstats: torch.Tensor, eps: float, gelu: bool = False) -> torch.Tensor:
    """out = act(LN(x) @ W^T + b) computed as a GEMM on the un-normalised x with the normalisation
    applied per row in the epilogue; ``rowstats`` is the (M, K/128, 2) fp32 table of per-128-column
    (sum, M2) partials of x's rows written by ``linear_res_stats``."""
    B, N, K = x.shape
    n_out = w_fold.shape[0]
    out = torch.empty((B, N, n_out), device=x.device, dtype=x.dtype)
    _lib.call("vt_gemm_bf16_ln", x.data_ptr(), K, w_fold.data_ptr(), K, out.data_ptr(), n_out,
              b_fold.data_ptr(), None, 0, B * N, n_out, K, 1 if gelu else 0, rowstats.data_ptr(),
              _lib.ptr(colsum), K, float(eps), None, _lib.stream_ptr(x))
    return out


def linear_res_stats(x: torch.Tensor, w_nk: torch.Tensor, bias32: torch.Tensor, residual: torch.Tensor,
                     stats_out: torch.Tensor) -> torch.Tensor:
    """out = x @ W^T + b + residual, also writing every output row's per-128-column (sum, M2) partials
    into the (M, N/128, 2) fp32 ``stats_out`` for the LayerNorm folded into the next GEMM."""
    B, N, K = x.shape
    n_out = w_nk.shape[0]
    out = torch.empty((B, N, n_out), device=x.device, dtype=x.dtype)
    _lib.call("vt_gemm_bf16_ln", x.data_ptr(), K, w_nk.data_ptr(), K, out.data_ptr(), n_out,
              bias32.data_ptr(), residual.data_ptr(), n_out, B * N, n_out, K, 0, None, None, 0, 0.0,
              stats_out.data_ptr(), _lib.stream_ptr(x))
    return out


def linear(x: torch.Tensor, w_nk: torch.Tensor, bias32: torch.Tensor, gelu: bool = False,
           residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epi(x @ w_nk^T + bias) over the flattened (B*N, K) activation.

    bf16 -> tcgen05 GEMM with bias / GELU / residual fused into the epilogue;
    fp32 -> exact FP32-pipe GEMM (bias / GELU fused, residual added by the add kernel).
    """
    B, N, K = x.shape
    n_out = w_nk.shape[0]
    out = torch.empty((B, N, n_out), device=x.device, dtype=x.dtype)
    M = B * N
    if M == 0:
        return out
    stream = _lib.stream_ptr(x)
    if x.dtype == torch.bfloat16 and K % 8 == 0 and n_out % 8 == 0:
        _lib.call("vt_gemm_bf16", x.data_ptr(), K, w_nk.data_ptr(), K, out.data_ptr(), n_out,
                  _lib.VT_BF16, bias32.data_ptr(), _lib.ptr(residual), n_out, M, n_out, K,
                  1 if gelu else 0, stream)
        return out

    if os.environ.get("VT_EXACT_FP32") == "1":      # FP32-pipe kernel, A/B comparisons only
        bias = bias32 if x.dtype == torch.float32 else bias32.to(x.dtype)
        code = _lib.dtype_code(x)
        step = 65535 * 64
        for m0 in range(0, M, step):
            mm = min(step, M - m0)
            es = x.element_size()
            _lib.call("vt_gemm_strided", x.data_ptr() + m0 * K * es, w_nk.data_ptr(),
                      out.data_ptr() + m0 * n_out * es, bias.data_ptr(), mm, n_out, K, 1, 1,
                      _lib.i64x4(0, 0, K, 1), _lib.i64x4(0, 0, 1, K), _lib.i64x4(0, 0, n_out, 1),
                      1.0, 1 if gelu else 0, code, stream)
        if residual is not None:
            _lib.call("vt_add", out.data_ptr(), residual.data_ptr(), out.data_ptr(), out.numel(), code, stream)
        return out

    # fp32 model (and bf16 layers whose widths are not multiples of 8): tensor cores through vt_bgemm —
    # fp32 operands split into bf16 pieces (kernels/bgemm.py), bias / GELU / residual fused
    pieces = 1 if x.dtype == torch.bfloat16 else bg.split_pieces()
    wp = split_weight(w_nk, pieces)
    bg.dense_rows(x, x.data_ptr(), M, K, K, wp, pieces, n_out, bias32, bg.ACT_GELU if gelu else bg.ACT_NONE, out,
                  out.data_ptr(), n_out, residual_ptr=_lib.ptr(residual))
    return out


# packed (split) forms of K-major weights, keyed on the packed tensor itself: the [N, K] matrices in the
# PackedMixin namespaces are rebuilt (new tensors) whenever a source parameter changes, so entries die
# with them
_split_cache = {}


def split_weight(w_nk: torch.Tensor, pieces: int) -> torch.Tensor:
    key = (id(w_nk), pieces)
    hit = _split_cache.get(key)
    if hit is not None and hit[0]() is w_nk:
        return hit[1]
    import weakref
    if len(_split_cache) > 1024:
        for k in [k for k, v in _split_cache.items() if v[0]() is None]:
            del _split_cache[k]
    value = bg.pack_weight_nk(w_nk, pieces)
    _split_cache[key] = (weakref.ref(w_nk), value)
    return value


def _two_launch_embed() -> bool:
    """Gather + token-mode GEMM (default) or round 1's single kernel (VT_PATCH_EMBED=1, A/B measurements)."""
    return os.environ.get("VT_PATCH_EMBED", "2") != "1"


def _embed_call(stats, b0, n_tok, dim, pk, pixels_ptr, pix_code, out_ptr, nb, C, S, P, stream, device):
    """Patch embedding of images b0 .. b0 + nb: ``vt_patch_embed_gemm`` (gather + 2-CTA GEMM in token mode), or the
    single-kernel ``vt_patch_embed[_stats]``; row statistics of those images go into ``stats`` when given."""
    s_ptr = None if stats is None else stats.data_ptr() + b0 * n_tok * (dim // STATS_COLS) * 2 * 4
    if _two_launch_embed():
        tok_pad = (n_tok + 31) // 32 * 32
        work = torch.empty((nb, tok_pad, pk.ldw), device=device, dtype=torch.bfloat16)
        _lib.call("vt_patch_embed_gemm", pixels_ptr, pix_code, pk.w.data_ptr(), pk.ldw, pk.bias32.data_ptr(),
                  pk.posb16.data_ptr(), out_ptr, s_ptr, work.data_ptr(), nb, C, S, P, dim, stream)
        return
    if stats is None:
        _lib.call("vt_patch_embed", pixels_ptr, pix_code, pk.w.data_ptr(), pk.ldw, pk.posb.data_ptr(), out_ptr,
                  _lib.VT_BF16, nb, C, S, P, dim, stream)
    else:
        _lib.call("vt_patch_embed_stats", pixels_ptr, pix_code, pk.w.data_ptr(), pk.ldw, pk.posb.data_ptr(), out_ptr,
                  _lib.VT_BF16, s_ptr, nb, C, S, P, dim, stream)


def patch_embed(emb, x: torch.Tensor, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pixels (B, C, S, S) -> token embeddings (B, n+1, D) including CLS and position embeddings.
    ``stats``: optional (B * (n+1), D/128, 2) fp32 buffer for the row statistics the LayerNorm folded
    into the first QKV GEMM consumes (bf16 tensor-core path only)."""
    pk = emb.packed()
    w = emb.projection.weight
    B, C, S, _ = x.shape
    P = emb.patch_size
    D = emb.hidden_dim
    n_tok = emb.num_patches + 1
    x = x.contiguous()
    out = torch.empty((B, n_tok, D), device=x.device, dtype=w.dtype)
    if B == 0:
        return out
    stream = _lib.stream_ptr(x)

    if w.dtype == torch.bfloat16 and D % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16):
        step = 65535 * 128 // emb.num_patches
        for b0 in range(0, B, step):
            nb = min(step, B - b0)
            _embed_call(stats, b0, n_tok, D, pk, x[b0:].data_ptr(), _lib.dtype_code(x), out[b0:].data_ptr(), nb, C, S, P,
                        stream, x.device)
        return out
    assert stats is None, "Row statistics come out of the bf16 tensor-core patch embedding only"

    # fp32 / odd-width path: im2col rows, tensor-core GEMM (fp32 operands split into bf16 pieces) straight
    # into rows 1..n of every image, then CLS / position embeddings
    from .kernels.patching import patching
    if x.dtype != w.dtype:
        x = x.to(w.dtype)
    patches = patching(x, P)
    n = emb.num_patches
    K = pk.K
    code = _lib.dtype_code(out)
    es = out.element_size()
    if os.environ.get("VT_EXACT_FP32") == "1":
        bias = pk.bias32 if w.dtype == torch.float32 else emb.projection.bias.detach().contiguous()
        for b0 in range(0, B, 32768):
            nb = min(32768, B - b0)
            _lib.call("vt_gemm_strided", patches[b0:].data_ptr(), pk.w.data_ptr(),
                      out[b0:].data_ptr() + D * es, bias.data_ptr(), n, D, K, nb, 1,
                      _lib.i64x4(n * K, 0, K, 1), _lib.i64x4(0, 0, 1, pk.ldw), _lib.i64x4(n_tok * D, 0, D, 1),
                      1.0, 0, code, stream)
    else:
        pieces = 1 if w.dtype == torch.bfloat16 else bg.split_pieces()
        wp = split_weight(pk.w, pieces)                    # [D, pieces * ldw]: pk.w rows are already padded to 8
        a = bg.pack(patches, patches.data_ptr(), n, K, B, 1, (n * K, 0, K, 1), pieces, pattern=0)
        kk = a.shape[2]
        bg.bgemm(a.data_ptr(), wp.data_ptr(), out, out.data_ptr() + D * es, n, D, kk, B, 1, (n * kk, 0, kk),
                 (0, 0, kk), (n_tok * D, 0, D), bias32=pk.bias32)
    _lib.call("vt_embed_finalize", out.data_ptr(), emb.position_embeddings.data_ptr(),
              emb.cls_token.data_ptr(), B, n_tok, D, code, stream)
    return out
