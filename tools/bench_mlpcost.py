"""MLPCost (milo/milo/linear_cost.py:154-301) features + cost + bonus combine on
device, measured like the env step.  One JSON line: rows/s through Engine.bonus_cost on device-resident rows,
the feature-net GEMMs' TFLOP/s against the measured tensor peak, and the oracle on the host cores.

    python tools/bench_mlpcost.py [--rows 40000] [--iters 50]
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
    ap.add_argument("--rows", type=int, default=40000)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    out_fd = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import bench as B
    from amp_extensions_b200 import MLPCost
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    expert = B.synth_expert(4096, 2)
    cost = MLPCost(expert, lambda_b=B.LAMBDA_B, seed=100, device=dev)
    eng = cost.engine()
    n = args.rows
    g = torch.Generator(device=dev).manual_seed(1)
    xs = [torch.randn(n, 452, device=dev, generator=g) for _ in range(4)]
    disc = torch.rand(n, device=dev, generator=g) * 0.5
    cost.fit_cost(xs[0][:1024].cpu())
    w = cost.w.to(dev)
    for i in range(3):
        eng.bonus_cost(xs[i % 4], disc, w, cost.lambda_b, 0.4)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.iters):
        eng.bonus_cost(xs[i % 4], disc, w, cost.lambda_b, 0.4)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / args.iters
    eng.profile_enable(True)
    for i in range(10):
        eng.bonus_cost(xs[i % 4], disc, w, cost.lambda_b, 0.4)
    prof = eng.profile_read(reset=True)
    eng.profile_enable(False)
    per = {k: v[0] / 10 for k, v in prof.items() if v[1] > 0}
    lin = cost._linears()
    flop_row = sum(2 * l.in_features * l.out_features for l in lin)
    gemm_ms = per.get("ensemble_gemm", 0.0) + per.get("rff_gemm", 0.0)
    peaks = B.measured_peaks()
    cpu = None
    if not args.skip_cpu:
        from oracle import milo_oracle as mo
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        oc = mo.MlpCostOracle(expert, lambda_b=B.LAMBDA_B, seed=100)
        oc.w = cost.w
        xc = xs[0][:4096].cpu()
        dc = disc[:4096].cpu()
        s, s2 = xc[:, :226], xc[:, 226:]
        oc.get_bonus_costs(s, None, dc, 0.4, next_states=s2)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 10.0:
            oc.get_bonus_costs(s, None, dc, 0.4, next_states=s2)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": 4096 / dt, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": f"{reps} x 4096 rows, oracle restatement of MLPCost.get_bonus_costs (discrepancy given)"}
    line = {"workload": f"MLPCost 452-2048-2048-1024 tanh/cos features + cost + bonus combine, {n} rows",
            "api": "amp_extensions_b200.MLPCost -> Engine.bonus_cost (device-resident rows)", "ms_per_call": ms,
            "rows_per_s": n / (ms * 1e-3), "kernels_ms": per,
            "roofline": {"bound": "tensor", "algorithmic_flop_per_row": flop_row,
                         "achieved": flop_row * n / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                         "peak": peaks["tensor_tflops"], "unit": "TFLOP/s"},
            "cpu_baseline": cpu}
    if line["roofline"]["achieved"]:
        line["roofline"]["frac"] = line["roofline"]["achieved"] / peaks["tensor_tflops"]
    out_fd.write(json.dumps(line) + "\n")
    out_fd.flush()


if __name__ == "__main__":
    main()
