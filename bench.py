#!/usr/bin/env python
"""Benchmark of the hot path: ViT forward, images/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--impl reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1).  A step = one forward
of the per-GPU batch (weak scaling: the per-GPU batch is fixed) + CLS pooling + the all-gather of
the pooled embeddings when N > 1 (one pool + peer-store kernel over NVLink, vit/parallel.py:PeerGather;
VT_PEER_GATHER=0 = vt_pool_cls + NCCL; `config.gather` says which ran).  Rank 0 prints ONE JSON line.

  value     device-resident inputs, CUDA-event timed, max over ranks
  e2e       same steps through the public API from PINNED HOST buffers: H2D of every step's pixels
            and D2H of its pooled embeddings inside the timed region (double-buffered copy stream)
  roofline  the tcgen05 GEMM kernel (95 % of the FLOPs): algorithmic FLOPs / CUDA-event time of its
            launches, against MEASURED_PEAKS.json; the events are recorded in a SECOND pass over the
            same K steps so that the timed region of `value` holds nothing but the steps
  cpu_baseline / --impl reference
            HuggingFace ViTModel fp32 on the box's host cores (the oracle and timing reference
            BASELINE.json names; the reference's own Triton kernels cannot run on a CPU).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "vit.triton_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CONFIGS = {
    # name: (arch, per-GPU batch, description)
    "c2": ("vit-b16-224", 256, "ViT-B/16@224 bf16 forward, batch 256 per GPU (BASELINE configs[1])"),
    "c3": ("vit-b16-384", 128, "ViT-B/16@384 (577 tokens) bf16 forward, batch 128 per GPU (BASELINE configs[2])"),
    "c4": ("vit-l16-224", 128, "ViT-L/16@224 bf16 forward, batch 128 per GPU (BASELINE configs[3] at 8 GPUs)"),
    "c5": ("vit-h14-224", 256, "ViT-H/14@224 bf16 forward, batch 256 per GPU (BASELINE configs[4])"),
}


def flops_per_image(a):
    """Algorithmic FLOPs (2*MAC) of GEMMs + attention per image."""
    D, L, F, P, S = a["hidden_size"], a["num_hidden_layers"], a["intermediate_size"], a["patch_size"], a["image_size"]
    n = (S // P) ** 2
    N = n + 1
    return 2 * n * 3 * P * P * D + L * (2 * N * D * 3 * D + 4 * N * N * D + 2 * N * D * D + 4 * N * D * F)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"burst": p.get("bf16_tflops"), "sustained": p.get("bf16_tflops_sustained"),
                "hbm": p.get("hbm_gbs"), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx, power = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx = float(parts[1])
                    power.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [c for c, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def hf_cpu_images_per_sec(arch, batch, budget_s=15.0):
    """HF ViTModel fp32 forward on the host cores, a bounded sample of about ``budget_s`` seconds:
    returns (img/s from the median forward, threads, list of forward times)."""
    import torch
    from oracle import hf_oracle
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(os.cpu_count() or 1)
    model = hf_oracle.build_hf(arch, seed=0)
    x = hf_oracle.make_input(arch, batch)
    times = []
    with torch.no_grad():
        t0 = time.perf_counter()
        hf_oracle.hf_forward(model, x)            # warm-up, also sizes the sample
        first = time.perf_counter() - t0
        reps = max(3, min(200, int(budget_s / max(first, 1e-3))))
        for _ in range(reps):
            t0 = time.perf_counter()
            hf_oracle.hf_forward(model, x)
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return batch / med, torch.get_num_threads(), times


def run_reference(args, arch, desc):
    """--impl reference: the CPU implementation of the path (HF ViTModel fp32), bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import hf_oracle
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and is
    # meant to use every host thread it can
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(os.cpu_count() or 1)
    sample_batch = 32
    model = hf_oracle.build_hf(arch, seed=0)
    x = hf_oracle.make_input(arch, sample_batch)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 3))):
            hf_oracle.hf_forward(model, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            hf_oracle.hf_forward(model, x)
        dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    value = sample_batch / (ms / 1e3)
    threads = torch.get_num_threads()
    sample = f"HF ViTModel fp32 CPU forward (oracle/hf_oracle.py), {sample_batch} images per step"
    line = {
        "impl": "reference", "metric": "images_per_sec", "value": value, "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "arch": arch, "sample_images_per_step": sample_batch},
        "cpu_baseline": {"value": value, "unit": "img/s", "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="bracket ONE extra step after the timed region with cudaProfilerStart/Stop "
                         "(ncu --profile-from-start off then captures exactly one forward)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    arch, batch, desc = CONFIGS[args.config]
    if args.batch:
        batch = args.batch
    if args.impl == "reference":
        run_reference(args, arch, desc)
        return

    import torch
    import torch.distributed as dist
    from vit import configs
    from vit.kernels import _lib
    from vit.parallel import DataParallelVIT
    from vit.vit import VIT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    a = configs.ARCHS[arch]
    # random-init weights of the named architecture (no network for checkpoints), same on all ranks
    torch.manual_seed(0)
    model = VIT(**configs.vit_kwargs(arch))
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() > 1:
                torch.nn.init.trunc_normal_(p, std=0.02)
            else:
                p.copy_(torch.randn_like(p) * 0.02)
        for m in model.modules():
            if type(m).__name__ == "LayerNormTriton":
                m.weight.add_(1.0)
    model = model.to(device=dev, dtype=torch.bfloat16).eval()
    dp = DataParallelVIT(model)
    global_batch = batch * world

    # synthetic pixels: several distinct input batches are rotated so no step re-reads a cached input
    S = a["image_size"]
    n_rot = 4
    g = torch.Generator().manual_seed(1234 + rank)
    host_inputs = [torch.randn((batch, 3, S, S), generator=g).to(torch.bfloat16).pin_memory() for _ in range(n_rot)]
    dev_inputs = [h.to(dev, non_blocking=True) for h in host_inputs]
    torch.cuda.synchronize()

    def step(x):
        return dp(x, global_batch)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ----------------------------------------------------------------- device-resident timing
    gemm_events = []

    def hook(name, before):
        if name in ("vt_gemm_bf16", "vt_gemm_bf16_ln"):   # the same kernel with / without the LayerNorm fold
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            gemm_events.append(ev)

    with torch.no_grad():
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()          # nvidia-smi needs ~0.5 s to produce its first sample
        for i in range(args.warmup):
            step(dev_inputs[i % n_rot])
        sync_all()
        # timed region: EXACTLY K steps, nothing but the steps between the two events
        launches0 = _lib.launch_count
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(args.steps):
            out = step(dev_inputs[i % n_rot])
        stop.record()
        sync_all()
        launches = _lib.launch_count - launches0
        # the same K steps again with a CUDA event before and after every GEMM launch (roofline of the
        # dominant kernel); kept out of the region above because 96 event records per step cost 2-4 %
        _lib.event_hook = hook
        h_start, h_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h_start.record()
        for i in range(args.steps):
            out = step(dev_inputs[i % n_rot])
        h_stop.record()
        _lib.event_hook = None
        sync_all()
        clocks = sampler.stop() if rank == 0 else None
    if args.profile_step:
        with torch.no_grad():
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            step(dev_inputs[0])
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
    ms_total = start.elapsed_time(stop)
    hooked_ms_total = h_start.elapsed_time(h_stop)
    gemm_ms = sum(gemm_events[i].elapsed_time(gemm_events[i + 1]) for i in range(0, len(gemm_events), 2))
    n_gemm = len(gemm_events) // 2

    # ----------------------------------------------------------------- end-to-end from pinned host memory
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(dev_inputs[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    pooled_host = torch.empty((global_batch, a["hidden_size"]), dtype=torch.bfloat16).pin_memory()

    def e2e_run(nsteps, hosts, devs):
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            devs[0].copy_(hosts[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(nsteps):
            slot = i & 1
            if i + 1 < nsteps:
                nxt = (i + 1) & 1
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(consumed[nxt])
                    devs[nxt].copy_(hosts[(i + 1) % n_rot], non_blocking=True)
                    ready[nxt].record(copy_stream)
            cur.wait_event(ready[slot])
            res = step(devs[slot])
            consumed[slot].record(cur)
            pooled_host.copy_(res, non_blocking=True)
        return res

    def e2e_time(hosts, devs):
        with torch.no_grad():
            e2e_run(3, hosts, devs)
            sync_all()
            e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_host0 = time.perf_counter()
            e_start.record()
            e2e_run(args.steps, hosts, devs)
            e_stop.record()
            sync_all()
            wall = (time.perf_counter() - t_host0) * 1e3
        return max(e_start.elapsed_time(e_stop), 0.0), wall

    e2e_ms, e2e_wall_ms = e2e_time(host_inputs, bufs)
    # same steps from RAW uint8 NHWC host pixels: the image processor's rescale + normalise run inside
    # the patch-embedding kernel (VIT.forward_uint8), so the H2D copy is one byte per pixel value
    g8 = torch.Generator().manual_seed(4321 + rank)
    host_u8 = [torch.randint(0, 256, (batch, S, S, 3), generator=g8, dtype=torch.uint8).pin_memory()
               for _ in range(n_rot)]
    bufs_u8 = [torch.empty((batch, S, S, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    e2e_u8_ms, _ = e2e_time(host_u8, bufs_u8)

    # ----------------------------------------------------------------- reduce over ranks
    stats = torch.tensor([ms_total, e2e_ms, gemm_ms, float(launches), e2e_u8_ms], device=dev, dtype=torch.float64)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_ms, e2e_u8_ms = mx[0].item(), mx[1].item(), mx[4].item()
        launches = int(sm[3].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = global_batch * args.steps / (ms_total / 1e3)
    e2e_value = global_batch * args.steps / (e2e_ms / 1e3)
    fpi = flops_per_image(a)
    peaks = measured_peaks()

    # GEMM FLOPs per step on this rank (QKV, proj, fc1, fc2 of every layer; patch-embed and attention
    # are separate kernels and not counted here)
    D, L, F = a["hidden_size"], a["num_hidden_layers"], a["intermediate_size"]
    N_tok = (S // a["patch_size"]) ** 2 + 1
    gemm_flops_step = L * batch * N_tok * (2 * D * 3 * D + 2 * D * D + 4 * D * F)
    gemm_tflops = gemm_flops_step * args.steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_dram_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "kernel": "gemm2_bf16_kernel (tcgen05 cta_group::2)", "bound": "tensor", "achieved": gemm_tflops,
        "peak": peaks["sustained"], "unit": "TFLOP/s",
        "frac": (gemm_tflops / peaks["sustained"]) if gemm_tflops else None,
        "frac_of_burst_peak": (gemm_tflops / peaks["burst"]) if gemm_tflops else None,
        "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
        "traffic": traffic, "launches_timed": n_gemm,
        "avg_launch_us": (gemm_ms / n_gemm * 1e3) if n_gemm else None,
        "algorithmic_flops_per_launch": gemm_flops_step / (4 * L),
        "share_of_step": gemm_ms / hooked_ms_total,
        "timed_in": "a second pass over the same K steps with a CUDA event around every GEMM launch "
                    f"({hooked_ms_total / args.steps:.3f} ms per step with the events)",
        "whole_forward_tflops": fpi * batch / (ms_per_step / 1e3) / 1e12,
        "whole_forward_frac_of_burst_peak": fpi * batch / (ms_per_step / 1e3) / 1e12 / peaks["burst"],
    }

    line = {
        "metric": "images_per_sec", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": desc, "arch": arch, "per_gpu_batch": batch, "global_batch": global_batch,
                   "tokens": N_tok, "parallelism": f"dp{world}" if world > 1 else "single",
                   "l2": f"{n_rot} distinct input batches rotated; per-step activation footprint > L2",
                   "weights": "random-init (trunc-normal 0.02), replicated",
                   "gather": ("pooled CLS embeddings, " + dp.gather_impl) if world > 1 else "none (one GPU)"},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "img/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": host_inputs[0].numel() * 2 * world,
                "d2h_bytes_per_step": pooled_host.numel() * 2 * world,
                "host_wall_ms_per_step": e2e_wall_ms / args.steps,
                "api": "DataParallelVIT(model)(pixels) from pinned host bf16 pixels, double-buffered H2D",
                "uint8_nhwc": {"value": global_batch * args.steps / (e2e_u8_ms / 1e3), "unit": "img/s",
                               "ms_per_step": e2e_u8_ms / args.steps,
                               "h2d_bytes_per_step": host_u8[0].numel() * world,
                               "api": "same, from pinned host RAW uint8 NHWC pixels (VIT.forward_uint8: "
                                      "rescale + normalise folded into the patch-embedding kernel)"}},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cpu_val, threads, times = hf_cpu_images_per_sec(arch, 32)
        line["cpu_baseline"] = {"value": cpu_val, "unit": "img/s", "cores": threads, "kind": "port",
                                "host_cpus": os.cpu_count(),
                                "sample": f"HF ViTModel fp32 CPU forward of {arch}, batch 32, median of {len(times)} "
                                          f"forwards after 1 warm-up ({sum(times):.1f} s of CPU work)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
