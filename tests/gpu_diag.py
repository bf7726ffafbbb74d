"""Developer diagnostics for a GPU box: runs each kernel family in its own subprocess (a trapped
kernel poisons its CUDA context, not the next family's) and prints compact error summaries.

    python tests/gpu_diag.py [family ...]      families: rowwise simt gemm attn patch model
"""
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vit.triton_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def rel(got, want):
    return ((got.float() - want.float()).norm() / want.float().norm().clamp_min(1e-12)).item()


def fam_rowwise():
    import torch
    import torch.nn.functional as F
    from vit.kernels import layernorm, add, softmax
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(4, 197, 768, device="cuda").to(dt)
        w = torch.randn(768, device="cuda").to(dt)
        b = torch.randn(768, device="cuda").to(dt)
        got = layernorm(x, w, b, 1e-12)
        want = F.layer_norm(x.float(), (768,), w.float(), b.float(), 1e-12)
        print("layernorm", dt, "max err", (got.float() - want).abs().max().item())
        print("add", dt, "exact", torch.equal(add(x, x), x + x))
        s = softmax(x)
        print("softmax", dt, "max err", (s.float() - torch.softmax(x.float(), -1)).abs().max().item())


def fam_simt():
    import torch
    from vit.kernels import matmul, matmul3
    a = torch.randn(4, 20, 30, device="cuda")
    b = torch.randn(30, 10, device="cuda")
    bias = torch.randn(10, device="cuda")
    got = matmul(a, b, bias, "gelu")
    want = torch.nn.functional.gelu(a @ b + bias)
    print("simt matmul gelu max err", (got - want).abs().max().item())
    a3 = torch.randn(4, 120, 760, device="cuda")
    b3 = torch.randn(4, 760, 500, device="cuda")
    print("matmul3 max err", (matmul3(a3, b3, True, 0.5) - 0.5 * (a3 @ b3)).abs().max().item())


def _gemm_case(M, K, N, act=None, res=False):
    import torch
    from vit.kernels import _lib
    torch.manual_seed(1)
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").bfloat16() if res else None
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    _lib.call("vt_gemm_bf16", x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, _lib.VT_BF16,
              bias.data_ptr(), None if r is None else r.data_ptr(), N, M, N, K, 1 if act else 0,
              _lib.stream_ptr(x))
    torch.cuda.synchronize()
    want = x.float() @ w.float().t() + bias
    if act:
        want = torch.nn.functional.gelu(want)
    if res:
        want = want + r.float()
    e = rel(out, want)
    print(f"gemm M={M} K={K} N={N} act={act} res={res}: rel err {e:.3e} max {(out.float()-want).abs().max().item():.3e}")
    if e > 1e-2:
        err = (out.float() - want).abs()
        bm = err.view(-1)[: (M // 8) * 8 * N].view(M // 8, 8, N).amax(dim=(1, 2)) if M >= 8 else err
        print("  rows-block(8) max err (first 16):", [round(v, 2) for v in bm[:16].tolist()])
        print("  col-block(8) max err (first 16):", [round(v, 2) for v in err[:, : (N // 8) * 8].view(M, N // 8, 8).amax(dim=(0, 2))[:16].tolist()])
        print("  out[0,:8]", out[0, :8].float().tolist())
        print("  want[0,:8]", want[0, :8].tolist())
    return e


def fam_gemm():
    import torch
    _gemm_case(128, 64, 128)
    _gemm_case(128, 64, 256)
    _gemm_case(128, 256, 256)
    _gemm_case(256, 768, 768)
    _gemm_case(197 * 8, 768, 2304)
    _gemm_case(197 * 8, 768, 3072, act="gelu")
    _gemm_case(197 * 8, 3072, 768, res=True)
    _gemm_case(300, 264, 40)
    # timing at C2 layer shapes
    from vit.kernels import _lib
    M = 256 * 197
    for (K, N, act, res) in ((768, 2304, 0, False), (768, 768, 0, True), (768, 768, 0, False), (768, 3072, 1, False), (768, 3072, 0, False), (3072, 768, 0, True), (3072, 768, 0, False)):
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        bias = torch.randn(N, device="cuda")
        r = torch.randn(M, N, device="cuda").bfloat16() if res else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        def run():
            _lib.call("vt_gemm_bf16", x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, _lib.VT_BF16,
                      bias.data_ptr(), None if r is None else r.data_ptr(), N, M, N, K, act, _lib.stream_ptr(x))
        for _ in range(3):
            run()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(50):
            run()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 50
        print(f"gemm C2 K={K} N={N} gelu={act} res={res}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s")


def _attn_ref(qkv, H):
    import torch
    B, N, D3 = qkv.shape
    D = D3 // 3
    dh = D // H
    q, k, v = qkv.float().split(D, dim=2)
    sh = lambda t: t.view(B, N, H, dh).transpose(1, 2)
    p = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) / math.sqrt(dh), dim=-1)
    return (p @ sh(v)).transpose(1, 2).reshape(B, N, D)


def fam_attn():
    import torch
    from vit.kernels import flash_attention
    torch.manual_seed(2)
    # structured probe: V = identity on the first 64 kv rows -> out[:, d] = P[:, d]
    N, H = 128, 1
    qkv = torch.zeros(1, N, 3 * 64, device="cuda")
    qkv[0, :, :64] = torch.randn(N, 64, device="cuda")
    qkv[0, :, 64:128] = torch.randn(N, 64, device="cuda")
    qkv[0, :64, 128:] = torch.eye(64, device="cuda")
    qkv = qkv.bfloat16()
    got = flash_attention(qkv, H)
    torch.cuda.synchronize()
    want = _attn_ref(qkv, H)
    print("attn probe (V=I) rel err", rel(got, want), "max", (got.float() - want).abs().max().item())
    if rel(got, want) > 2e-2:
        print("  got[0,:8]", got[0, 0, :8].float().tolist())
        print("  want[0,:8]", want[0, 0, :8].tolist())
        print("  got[1,:8]", got[0, 1, :8].float().tolist())
        print("  want[1,:8]", want[0, 1, :8].tolist())
        print("  row sums got", got[0, :4].float().sum(-1).tolist(), "want", want[0, :4].sum(-1).tolist())
    for (B, H, N) in ((1, 1, 16), (1, 1, 64), (1, 1, 128), (1, 2, 17), (2, 12, 197), (1, 16, 257), (2, 12, 577)):
        qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16()
        got = flash_attention(qkv, H)
        torch.cuda.synchronize()
        want = _attn_ref(qkv, H)
        print(f"attn B={B} H={H} N={N}: rel err {rel(got, want):.3e} max {(got.float()-want).abs().max().item():.3e} finite {bool(torch.isfinite(got.float()).all())}")
    for (B, H, N) in ((1, 1, 16), (1, 2, 64), (1, 1, 128), (2, 16, 257), (1, 3, 500)):
        qkv = torch.randn(B, N, 3 * H * 80, device="cuda").bfloat16()
        got = flash_attention(qkv, H)
        torch.cuda.synchronize()
        want = _attn_ref(qkv, H)
        print(f"attn dh80 B={B} H={H} N={N}: rel err {rel(got, want):.3e} max {(got.float()-want).abs().max().item():.3e} finite {bool(torch.isfinite(got.float()).all())}")
        if rel(got, want) > 2e-2:
            e = (got.float() - want).abs()[0]
            hd = e.view(N, H, 80)
            print("   err by head-dim column block (16 wide):", [round(hd[:, :, i:i+16].max().item(), 3) for i in range(0, 80, 16)])
            print("   err by row block (32):", [round(e[i:i+32].max().item(), 3) for i in range(0, min(N, 256), 32)])
    for (B, H, N) in ((256, 12, 197), (128, 12, 577)):
        qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16()
        for _ in range(3):
            flash_attention(qkv, H)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            flash_attention(qkv, H)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        print(f"attn B={B} H={H} N={N}: {ms*1e3:.1f} us  {4*B*H*N*N*64/ms/1e9:.0f} TFLOP/s")


def fam_patch():
    import torch
    import torch.nn.functional as F
    from vit.vit import Embeddings
    torch.manual_seed(3)
    for (S, P, D, B) in ((64, 16, 128, 2), (224, 16, 768, 3), (224, 14, 1280, 2)):
        n = (S // P) ** 2
        emb = Embeddings(P, n, 3 * P * P, D).cuda()
        with torch.no_grad():
            for p_ in emb.parameters():
                p_.copy_(torch.randn_like(p_) * 0.05)
        emb = emb.to(torch.bfloat16)
        x = torch.randn(B, 3, S, S, device="cuda").bfloat16()
        got = emb(x)
        torch.cuda.synchronize()
        tok = F.conv2d(x.float(), emb.projection.weight.float(), emb.projection.bias.float(), stride=P).flatten(2).transpose(1, 2)
        want = torch.cat([emb.cls_token.float().expand(B, -1, -1), tok], 1) + emb.position_embeddings.float()
        print(f"patch S={S} P={P} D={D}: rel err {rel(got, want):.3e} max {(got.float()-want).abs().max().item():.3e}")
        if rel(got, want) > 1e-2:
            print("  cls err", (got[:, 0].float() - want[:, 0]).abs().max().item(), "tok err", (got[:, 1:].float() - want[:, 1:]).abs().max().item())


def fam_model():
    import torch
    from oracle import hf_oracle
    from vit.utils import transfer_pretrained_weights
    from vit.vit import VIT
    for arch, dt, b in (("tiny-b", torch.float32, 2), ("tiny-b", torch.bfloat16, 2), ("vit-b16-224", torch.float32, 1), ("vit-b16-224", torch.bfloat16, 2)):
        hf = hf_oracle.build_hf(arch, seed=0)
        m = VIT(**hf_oracle.vit_kwargs(arch))
        transfer_pretrained_weights(hf, m, verbose=False)
        m = m.to("cuda", dt)
        x = hf_oracle.make_input(arch, b)
        want = hf_oracle.hf_forward(hf, x)
        with torch.no_grad():
            got = m(x.to("cuda", dt)).float().cpu()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
        print(f"model {arch} {dt} b={b}: max-abs {(got-want).abs().max().item():.3e} cosine {cos:.6f}")
    # C2 timing
    arch = "vit-b16-224"
    m = VIT(**hf_oracle.vit_kwargs(arch)).to("cuda", torch.bfloat16)
    with torch.no_grad():
        for p_ in m.parameters():
            p_.copy_(torch.randn_like(p_) * 0.02)
    x = torch.randn(256, 3, 224, 224, device="cuda").bfloat16()
    with torch.no_grad():
        for _ in range(3):
            m(x)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            m(x)
        e.record()
        torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"C2 forward b256: {ms:.2f} ms  {256/ms*1e3:.0f} img/s  {256*35.126/ms:.0f} TFLOP/s")


FAMILIES = {"rowwise": fam_rowwise, "simt": fam_simt, "gemm": fam_gemm, "attn": fam_attn, "patch": fam_patch, "model": fam_model}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        FAMILIES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(FAMILIES)
    for name in names:
        t0 = time.time()
        print(f"===== {name}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name], timeout=300,
                               capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print(f"[{name}] exit {r.returncode}\n{r.stderr[-3000:]}")
        except subprocess.TimeoutExpired:
            print(f"[{name}] TIMEOUT")
        print(f"===== {name} done in {time.time()-t0:.1f}s", flush=True)
