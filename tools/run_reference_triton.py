"""Run the UNMODIFIED reference (cmeraki/vit.triton, copied to baseline/_ref by oracle/fetch_reference.sh) on
this GPU: its own VIT module, its own Triton kernels, weights moved by its own loader.

    python tools/run_reference_triton.py <in.pt> <out.pt> [batch sizes to time ...]

in.pt  : {"state_dict": HF ViTModel state-dict (ViT-B/16@224, add_pooling_layer=False), "input": (B,3,224,224) fp32}
out.pt : {"output": reference forward of the input (fp32, CPU), "timings_ms": {batch: median ms}, "error": str | None,
          "triton": version}

Runs as its OWN process with baseline/_ref first on sys.path: the reference package is called ``vit`` like
ours, so the two cannot share an interpreter.  TEST / MEASUREMENT INFRASTRUCTURE: nothing here is imported by
the product."""
import os
import statistics
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    batches = [int(b) for b in sys.argv[3:]]
    import torch
    result = {"output": None, "timings_ms": {}, "error": None, "triton": None}
    try:
        if not os.path.isdir(os.path.join(REF, "vit")):
            raise RuntimeError("baseline/_ref/vit missing: run oracle/fetch_reference.sh in the build container")
        sys.path.insert(0, REF)
        import triton
        result["triton"] = triton.__version__
        from transformers import ViTConfig, ViTModel
        from vit.vit import VIT                      # the REFERENCE's module (baseline/_ref/vit/vit.py:203)
        from vit.utils import transfer_pretrained_weights
        blob = torch.load(src)
        hf = ViTModel(ViTConfig(), add_pooling_layer=False).eval()
        hf.load_state_dict(blob["state_dict"])
        hf = hf.to("cuda:0", torch.float32)
        model = VIT(height=224, width=224, channels=3, patch_size=16, hidden_dim=768, num_heads=12, num_layers=12)
        model.to(device="cuda:0", dtype=torch.float32)
        model = transfer_pretrained_weights(pretrained_model=hf, custom_model=model)
        x = blob["input"].to("cuda:0", torch.float32)
        with torch.no_grad():
            out = model(x)
            torch.cuda.synchronize()
            result["output"] = out.float().cpu()
            for b in batches:
                xb = torch.randn(b, 3, 224, 224, device="cuda:0")
                for _ in range(3):
                    model(xb)
                torch.cuda.synchronize()
                times = []
                for _ in range(10):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    model(xb)
                    e1.record()
                    torch.cuda.synchronize()
                    times.append(e0.elapsed_time(e1))
                result["timings_ms"][b] = statistics.median(times)
    except Exception:
        result["error"] = traceback.format_exc()[-3000:]
    torch.save(result, dst)
    print("reference run:", "ok" if result["error"] is None else "FAILED", result["timings_ms"], flush=True)
    if result["error"]:
        print(result["error"], flush=True)


if __name__ == "__main__":
    main()
