"""End-to-end XFMamba workloads for bench.py (`--workload xfmamba_*`): BASELINE.json configs 3, 4, 5.

    python bench.py --workload xfmamba_s_infer  --batch 64            # config 3: XFMamba-S, 13 classes, inference
    python bench.py --workload xfmamba_b_train  --batch 32 --gpus N   # config 4: XFMamba-B, training step, DDP + NCCL
    python bench.py --workload xfmamba_b_hires  --batch 8             # config 5: XFMamba-B, 512x512 inference

Synthetic two-view images (randn, in_channels=1) and random-init weights of the published architectures.  `value` is
pairs/s with the images resident in HBM; `e2e` copies every step's images from pinned host memory and reads the logits
(inference) or the loss (training) back.  Inference runs under torch.no_grad and, unless --no-graph, from a CUDA graph
(the scan library neither allocates nor synchronises, so the whole forward captures).
"""
from __future__ import annotations

import json
import os
import time

WORKLOADS = {
    "xfmamba_t_infer": dict(type="tiny", outputs=2, img=224, train=False),
    "xfmamba_s_infer": dict(type="small", outputs=13, img=224, train=False),
    "xfmamba_b_infer": dict(type="base", outputs=2, img=224, train=False),
    "xfmamba_b_train": dict(type="base", outputs=2, img=224, train=True),
    "xfmamba_s_train": dict(type="small", outputs=13, img=224, train=True),
    "xfmamba_b_hires": dict(type="base", outputs=2, img=512, train=False),
}


def train_leg(workload, batch, steps, warmup, dev, world, rank, local, amp=False):
    """BASELINE config 4 as a leg of the default bench.py run: XFMamba-B training step (fwd + bwd + gradient all-reduce +
    Adam) on `batch` synthetic two-view pairs per GPU; the process group (if any) is already initialised.  Returns the
    `model_train` object of the JSON line.  `allreduce_exposed_ms` = step time with DDP's bucketed NCCL all-reduce minus the
    step time of the same model under `no_sync()` (no collective at all), max over ranks."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from xfmamba_b200 import _lib
    from xfmamba_b200.model import TwoViewXFMamba

    w = WORKLOADS[workload]
    torch.manual_seed(0)
    model = TwoViewXFMamba(outputs=w["outputs"], type=w["type"]).to(dev)
    nparam = sum(p.numel() for p in model.parameters())
    gen = torch.Generator(device="cpu").manual_seed(rank)
    img = w["img"]
    xa = torch.randn(batch, 1, img, img, generator=gen).to(dev)
    xb = torch.randn(batch, 1, img, img, generator=gen).to(dev)
    yt = torch.randint(0, w["outputs"], (batch,), generator=gen).to(dev)
    model.train()
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True) if world > 1 else model
    opt = torch.optim.Adam(ddp.parameters(), lr=1e-4, weight_decay=1e-5)          # reference 1_train_model.py:135-141

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            loss = F.cross_entropy(ddp(xa, xb).float(), yt)
        loss.backward()
        opt.step()
        return loss

    def timed(n, nosync=False):
        import contextlib
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = _lib.launch_count()
        e0.record()
        with (ddp.no_sync() if (nosync and world > 1) else contextlib.nullcontext()):
            for _ in range(n):
                step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, (_lib.launch_count() - before) // n

    for _ in range(warmup):
        step()
    ms, launches = timed(steps)
    out = {"workload": workload, "model": f"XFMamba-{w['type']}", "params_m": round(nparam / 1e6, 2), "pairs_per_gpu": batch,
           "global_pairs": batch * world, "steps": steps, "ms_per_step": ms, "pairs_per_s": batch * world / (ms / 1e3),
           "grad_bytes_allreduced": nparam * 4 if world > 1 else 0, "xfscan_launches_per_step": launches,
           "dtype": "bf16 autocast" if amp else "f32", "mode": "fwd + bwd + DDP/NCCL gradient all-reduce + Adam"}
    if world > 1:
        timed(2, nosync=True)                          # untimed: the first no_sync steps re-allocate the gradients outside DDP's buckets
        ms_ns, _ = timed(max(3, steps // 2), nosync=True)
        out["ms_per_step_no_allreduce"] = ms_ns
        out["allreduce_exposed_ms"] = max(0.0, ms - ms_ns)
    else:
        out["allreduce_exposed_ms"] = 0.0
    del ddp, model, opt, xa, xb, yt
    torch.cuda.empty_cache()
    return out


def run(args, ClockSampler):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from xfmamba_b200 import _lib
    from xfmamba_b200.model import TwoViewXFMamba

    w = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    batch, img = args.batch, w["img"]
    model = TwoViewXFMamba(outputs=w["outputs"], type=w["type"]).to(dev)
    nparam = sum(p.numel() for p in model.parameters())
    amp = args.dtype == "bf16"
    gen = torch.Generator(device="cpu").manual_seed(rank)
    host_a = torch.randn(batch, 1, img, img, generator=gen).pin_memory()
    host_b = torch.randn(batch, 1, img, img, generator=gen).pin_memory()
    host_y = torch.randint(0, w["outputs"], (batch,), generator=gen).pin_memory()
    xa, xb, yt = host_a.to(dev), host_b.to(dev), host_y.to(dev)
    h2d = host_a.numel() * 4 * 2 + (host_y.numel() * 8 if w["train"] else 0)

    if w["train"]:
        model.train()
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True) if world > 1 else model
        opt = torch.optim.Adam(ddp.parameters(), lr=1e-4, weight_decay=1e-5)      # reference 1_train_model.py:135-141

        def step(a, b, y):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                loss = F.cross_entropy(ddp(a, b).float(), y)
            loss.backward()
            opt.step()
            return loss
        d2h = 4
    else:
        model.eval()
        static_out = None
        graph = None
        graph_launches = 0

        def fwd(a, b):
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                return model(a, b)
        if not args.no_graph:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    fwd(xa, xb)
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            cap0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                static_out = fwd(xa, xb)
            graph_launches = _lib.launch_count() - cap0        # xfscan kernels recorded in the graph = launched by every replay

        def step(a, b, y):
            if graph is not None:
                if a is not xa:
                    xa.copy_(a, non_blocking=True)
                    xb.copy_(b, non_blocking=True)
                graph.replay()
                return static_out
            return fwd(a, b)
        d2h = batch * w["outputs"] * 4

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        step(xa, xb, yt)
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    before = _lib.launch_count()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        step(xa, xb, yt)
    e1.record()
    sync_all()
    launches = _lib.launch_count() - before
    if not w["train"] and graph is not None:
        launches = graph_launches * args.steps           # replays launch the captured kernels without passing through the C ABI
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)

    # e2e: images start in pinned host memory every step, result is read back
    res_host = torch.empty(1 if w["train"] else batch * w["outputs"], dtype=torch.float32).pin_memory()
    da, db, dy = torch.empty_like(xa), torch.empty_like(xb), torch.empty_like(yt)

    def e2e_step():
        da.copy_(host_a, non_blocking=True)
        db.copy_(host_b, non_blocking=True)
        if w["train"]:
            dy.copy_(host_y, non_blocking=True)
        out = step(da, db, dy)
        res_host.copy_(out.detach().float().reshape(-1), non_blocking=True)
    for _ in range(2):
        e2e_step()
    sync_all()
    e2, e3 = ev(), ev()
    e2.record()
    for _ in range(args.steps):
        e2e_step()
    e3.record()
    sync_all()
    ms_e2e = e2.elapsed_time(e3)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    pairs = world * batch
    line = {
        "metric": "two_view_pairs_per_sec", "value": pairs * args.steps / (ms / 1e3), "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "model": f"XFMamba-{w['type']}", "params_m": round(nparam / 1e6, 2), "image": img,
                   "pairs_per_gpu": batch, "global_pairs": pairs, "mode": "train (fwd+bwd+allreduce+Adam)" if w["train"] else
                   ("inference, CUDA graph" if not args.no_graph else "inference, eager"), "parallelism": f"dp{world}"},
        "e2e": {"value": pairs * args.steps / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
