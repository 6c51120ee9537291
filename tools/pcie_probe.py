"""Host<->device copy bandwidth with every rank copying at once (explains bench.py's e2e at N GPUs).

    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py [--mb 36] [--no-bind]

Each rank times D2H, H2D and both together between a device buffer and pinned host memory (after binding to the
cores NVML reports closest to its GPU unless --no-bind); rank 0 prints per-rank and aggregate GB/s as one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=36)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--no-bind", action="store_true")
    args = ap.parse_args()
    out_fd = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from amp_extensions_b200 import parallel
    cpus = None if args.no_bind else parallel.bind_host_to_gpu(local)
    n = args.mb * (1 << 20) // 4
    d_a, d_b = torch.empty(n, device=dev), torch.empty(n, device=dev)
    h_a, h_b = torch.empty(n, pin_memory=True), torch.empty(n, pin_memory=True)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(kind):
        def once():
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s1):
                    h_a.copy_(d_a, non_blocking=True)
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s2):
                    d_b.copy_(h_b, non_blocking=True)
        for _ in range(3):
            once()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            once()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        gbs = args.mb * (1 << 20) * args.iters * (2 if kind == "both" else 1) / dt / 1e9
        t = torch.tensor([gbs], device=dev, dtype=torch.float64)
        allv = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(allv, t)
        else:
            allv = [t]
        return [float(x.item()) for x in allv]

    res = {k: run(k) for k in ("d2h", "h2d", "both")}
    if rank == 0:
        out_fd.write(json.dumps({"n_gpus": world, "mb_per_copy": args.mb, "bound": cpus is not None,
                                 "per_rank_gbs": {k: [round(x, 1) for x in v] for k, v in res.items()},
                                 "aggregate_gbs": {k: round(sum(v), 1) for k, v in res.items()}}) + "\n")
        out_fd.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
