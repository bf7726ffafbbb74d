"""Worker of test_peer_gather_matches_nccl_world2 / tools: run under torchrun or mp.spawn with NCCL.

Every rank builds the same tiny VIT, runs DataParallelVIT twice per step — through the fused
pool + peer-store kernel (vt_pool_cls_allgather over symmetric memory) and through NCCL — and
requires bit equality, for several steps (the flag counters and the gather buffers advance); the gathered
rows must also equal the single-process forward of the concatenated batch, in the
synchronous, the pipelined (submit / flush) and the CUDA-graph-replayed form."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "vit.triton_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch
import torch.distributed as dist


def run(rank: int, world: int, port: int, steps: int = 5) -> None:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from vit import configs
        from vit.parallel import DataParallelVIT
        from vit.vit import VIT
        torch.manual_seed(0)
        model = VIT(**configs.vit_kwargs("tiny-b")).to("cuda", torch.bfloat16)
        with torch.no_grad():
            for p_ in model.parameters():
                p_.copy_(torch.randn_like(p_) * 0.05)
        fused = DataParallelVIT(model, peer_gather=True)
        plain = DataParallelVIT(model, peer_gather=False)
        size = configs.ARCHS["tiny-b"]["image_size"]

        def shard(step, r, n=6):
            g = torch.Generator(device="cuda").manual_seed(100 * step + r)
            return torch.randn(n, 3, size, size, device="cuda", generator=g).bfloat16()

        with torch.no_grad():
            for step in range(steps):
                x = shard(step, rank)
                a = fused(x).clone()
                b = plain(x)
                assert fused.gather_impl == "peer-store kernel" and plain.gather_impl == "torch.distributed"
                assert a.shape == b.shape == (6 * world, b.shape[1])
                assert torch.equal(a, b), f"rank {rank} step {step}: peer-store gather differs from NCCL"
                # gathered == single-process forward of the concatenated batch, bit for bit
                whole = model.pooled(torch.cat([shard(step, r) for r in range(world)], dim=0))
                assert torch.equal(a, whole), f"rank {rank} step {step}: gathered rows differ from the single-process forward"
            # pipelined form: submit() returns the PREVIOUS step's gathered rows, flush() the last
            wants = []
            got = []
            for step in range(steps, steps + 4):
                wants.append(model.pooled(torch.cat([shard(step, r) for r in range(world)], dim=0)))
                prev = fused.submit(shard(step, rank))
                if prev is not None:
                    got.append(prev.clone())
            got.append(fused.flush().clone())
            assert fused.flush() is None
            assert len(got) == len(wants) and all(torch.equal(a, b) for a, b in zip(got, wants)), \
                f"rank {rank}: pipelined gather differs from the single-process forward"
            # back to the synchronous form, captured into a CUDA graph: every replay is the next step
            static_x = shard(0, rank).clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fused(static_x)                                           # warm-up (plans, packed weights)
            torch.cuda.current_stream().wait_stream(side)
            dist.barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = fused(static_x)
            for step in range(20, 24):
                static_x.copy_(shard(step, rank))
                graph.replay()
                whole = model.pooled(torch.cat([shard(step, r) for r in range(world)], dim=0))
                assert torch.equal(static_out, whole), f"rank {rank}: graph replay {step} differs"
            # HF pooler output (tanh(dense(CLS))) through the same kernel
            torch.manual_seed(2)
            pm = VIT(**configs.vit_kwargs("tiny-b"), add_pooling_layer=True).to("cuda", torch.bfloat16)
            for p_ in pm.parameters():
                p_.copy_(torch.randn_like(p_) * 0.05)
            fused_p = DataParallelVIT(pm, peer_gather=True, output="pooler")
            for step in range(2):
                a = fused_p(shard(step, rank, 4)).clone()
                whole = pm.pooler_output(torch.cat([shard(step, r, 4) for r in range(world)], dim=0))
                assert fused_p.gather_impl == "peer-store kernel" and torch.equal(a, whole)
            # class logits through the same kernel ((B, 1, num_labels) "hidden states")
            torch.manual_seed(1)
            clf = VIT(**configs.vit_kwargs("tiny-b"), num_labels=40).to("cuda", torch.bfloat16)
            for p_ in clf.parameters():
                p_.copy_(torch.randn_like(p_) * 0.05)
            fused_l = DataParallelVIT(clf, peer_gather=True, output="logits")
            plain_l = DataParallelVIT(clf, peer_gather=False, output="logits")
            for step in range(3):
                g = torch.Generator(device="cuda").manual_seed(7 * step + rank)
                x = torch.randn(4, 3, size, size, device="cuda", generator=g).bfloat16()
                a = fused_l(x).clone()
                b = plain_l(x)
                assert fused_l.gather_impl == "peer-store kernel" and a.shape == b.shape == (4 * world, 40)
                assert torch.equal(a, b), f"rank {rank} step {step}: gathered logits differ from NCCL"
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            print("peer gather == nccl gather on", world, "ranks,", steps, "steps", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    run(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("MASTER_PORT", "29511")))
