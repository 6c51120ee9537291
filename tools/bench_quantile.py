"""Latency of parallel.global_quantile on one GPU (< 1 ms per call at 1 M samples).

    python tools/bench_quantile.py [--n 1048576] [--reps 50]

Prints one JSON line: ms per call (host wall around the call, device idle before it), the result against
torch.quantile, and the library launches per call.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    from amp_extensions_b200 import _lib, parallel
    from amp_extensions_b200.engine import Engine
    eng = Engine(8, 2, 2, [16], precision="fp16")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(args.n, device="cuda", generator=g) ** 2
    out = {}
    for q in (0.1, 0.5, 0.9):
        ref = float(torch.quantile(x[:min(args.n, 1 << 24)].double(), q))
        got = parallel.global_quantile(x, q, engine=eng)
        out[f"q{q}"] = {"got": got, "torch": ref, "abs_err": abs(got - ref)}
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        parallel.global_quantile(x, 0.9, engine=eng)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.reps
    print(json.dumps({"n": args.n, "ms_per_call": dt * 1e3, "library_launches_per_call": (_lib.launch_count() - l0) / args.reps,
                      "check": out}))


if __name__ == "__main__":
    main()
