"""TEST INFRASTRUCTURE — torch restatement of the pack-time LayerNorm fold (csrc/ln_fold.cu, ``vt_ln_fold``).

Runs on any device (the CPU tests use it to check the arithmetic of the fold: zero-sum rows, one extra
ulp per element at most, same accuracy as the column-sum form); the GPU tests compare the kernel with
it.  The product path never imports this module: ``vit/packing.py:fold_layernorm`` calls the kernel.
The algorithm folded is the reference's LayerNorm (vit/kernels/layernorm.py:51-85) followed by its
dense layer (vit/kernels/matmul.py:73-108).
"""
import torch


def fold_layernorm_colsum(w_nk: torch.Tensor, bias32: torch.Tensor, ln) -> tuple:
    """Fold y = LN(x) into the dense layer that consumes it:  LN(x) @ W^T + b
         = rstd * (x @ (W * gamma)^T) - rstd * mean * colsum(W * gamma) + (b + W @ beta).
    Returns (W * gamma in the weight dtype, b + W @ beta in fp32, row sums of the ROUNDED folded weight)."""
    w32 = w_nk.float()
    w_fold = (w32 * ln.weight.detach().float()[None, :]).to(w_nk.dtype).contiguous()
    b_fold = (bias32 + w32 @ ln.bias.detach().float()).contiguous()
    colsum = w_fold.float().sum(dim=1).contiguous()
    return w_fold, b_fold, colsum


def fold_layernorm_zero_sum(w_nk: torch.Tensor, bias32: torch.Tensor, ln, passes: int = 6) -> tuple:
    """The same fold with the mean term moved INTO the weights: every row of W * gamma is shifted by
    its own mean, so that  x @ W'^T = x @ (W * gamma)^T - mean(x) * colsum(W * gamma)  comes out of the
    GEMM itself and the epilogue is just  rstd * acc + (b + W @ beta)  — no column-sum operand, one FMA
    per element less.  What is left of the mean term is  rstd * mean * sum_k(rounded W'[n, k]); the
    rounding residue of each row (~3e-3 after plain bf16 rounding of a 768-wide row) is cancelled by
    rounding a few elements the OTHER way — those closest to a rounding tie first, so the move costs
    almost nothing in accuracy — which leaves |sum_k W'| at the 1e-6 level.  Returns (W', b + W @ beta)."""
    w32 = w_nk.float()
    wg = w32 * ln.weight.detach().float()[None, :]
    b_fold = (bias32 + w32 @ ln.bias.detach().float()).contiguous()
    exact = wg - wg.mean(dim=1, keepdim=True)
    wz = exact.to(w_nk.dtype)
    if w_nk.dtype == torch.bfloat16:
        moved = torch.zeros(wz.shape, dtype=torch.bool, device=wz.device)   # every element moves at most once
        for _ in range(passes):
            f = wz.float()
            resid = f.double().sum(dim=1).float()[:, None]                   # (N, 1): what has to go
            # one bf16 ulp of every element: 2^(exponent - 7), built from the exponent field (integer ops only)
            expo = (f.view(torch.int32) >> 23) & 0xFF
            ulp = ((expo - 7).clamp_min(1) << 23).view(torch.float32)
            err = exact - f                                                  # rounding error so far
            ok = ~moved & (ulp <= resid.abs())
            # added squared error per unit of residue removed: small for elements that were rounded
            # the wrong way by almost half an ulp, large for those rounded the right way already
            helps = err * resid < 0
            price = torch.where(ok, ulp + torch.where(helps, -2.0, 2.0) * err.abs(), torch.full_like(ulp, float("inf")))
            order = price.argsort(dim=1)
            step = torch.where(ok, ulp, torch.zeros_like(ulp)).gather(1, order)
            take = (step.cumsum(dim=1) <= resid.abs()) & (step > 0)          # cheapest prefix that fits
            delta = torch.zeros_like(f).scatter(1, order, torch.where(take, step, torch.zeros_like(step)))
            f = f - torch.sign(resid) * delta                                # exact: one ulp of each element
            moved |= delta > 0
            wz = f.to(torch.bfloat16)
    return wz.contiguous(), b_fold
