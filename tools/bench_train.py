"""the ensemble training step on device (reference milo/milo/dynamics.py:236-250),
measured like the env step.  One JSON line: optimiser steps/s and transitions/s for the 4 x (512 x 4) humanoid3d
ensemble at the reference's batch size (256 rows per member, launch-latency bound) and at a large batch (tensor
bound), algorithmic training FLOP/s against the measured tensor peak, and the torch-autograd restatement of the
reference's train_step on the host cores.

    python tools/bench_train.py [--iters 50]
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
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    out_fd = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import bench as B
    from amp_extensions_b200.engine import Engine
    from oracle import milo_oracle as mo
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    S, A, N, hidden = B.S_DIM, B.A_DIM, B.N_MODELS, B.HIDDEN
    ws, bs = mo.init_ensemble(S, A, hidden, N, dense_connect=True, base_seed=100)
    s, a, s2 = B.synth_dataset(8192, 0)
    tf = mo.get_transformations(s, a, s2)
    flop_fwd = B.flops_per_env_step(N, hidden) / N            # per member per row
    res = {}
    for rows in (256, 16384):
        eng = Engine(S, A, N, hidden, dense_connect=True, transform=True, precision="tf32", device=dev)
        eng.train_init(rows, optim="sgd", lr=1e-4, momentum=0.9)
        eng.load_ensemble(ws, bs, tf)
        g = torch.Generator(device=dev).manual_seed(1)
        idx = torch.randint(0, 8192, (N, rows), device=dev, generator=g)
        sd, ad, nd = s.to(dev)[idx].contiguous(), a.to(dev)[idx].contiguous(), s2.to(dev)[idx].contiguous()
        for _ in range(3):
            eng.train_step(sd, ad, nd, grad_clip=1.0)
        torch.cuda.synchronize(dev)
        iters = args.iters if rows == 256 else max(5, args.iters // 5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            eng.train_step(sd, ad, nd, grad_clip=1.0)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / iters
        # forward + dgrad (hidden inputs only) + wgrad ~ 3 x forward FLOPs, the usual count for an MLP
        tflops = 3 * flop_fwd * rows * N / (ms * 1e-3) / 1e12
        res[f"rows{rows}"] = {"rows_per_member": rows, "ms_per_step": ms, "steps_per_s": 1e3 / ms,
                              "transitions_per_s": rows * N / (ms * 1e-3), "algorithmic_tflops": tflops}
        for _ in range(3):
            eng.train_step_graph(sd, ad, nd, grad_clip=1.0)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(iters):
            eng.train_step_graph(sd, ad, nd, grad_clip=1.0)
        e1.record()
        torch.cuda.synchronize(dev)
        msg = e0.elapsed_time(e1) / iters
        res[f"rows{rows}"]["graph"] = {"ms_per_step": msg, "transitions_per_s": rows * N / (msg * 1e-3),
                                      "algorithmic_tflops": 3 * flop_fwd * rows * N / (msg * 1e-3) / 1e12}
        eng.close()
    cpu = None
    if not args.skip_cpu:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        oracles = [mo.TrainOracle(ws[k], bs[k], tf, True, "relu") for k in range(N)]
        bi = torch.arange(256)
        for o in oracles:
            o.train_step(1.0, s[bi], a[bi], s2[bi])
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 10.0:
            for o in oracles:          # the reference trains the members one after the other
                o.train_step(1.0, s[bi], a[bi], s2[bi])
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": 256 * N / dt, "unit": "transitions/s", "cores": threads, "kind": "port",
               "sample": f"{reps} x (4 members x 256 rows), torch-autograd restatement of DynamicsModel.train_step"}
    peaks = B.measured_peaks()
    line = {"workload": "ensemble training step, 4 x (512 x 4) dense-connect, SGD-Nesterov + clip_grad_norm, tf32 operands",
            "api": "amp_extensions_b200.engine.Engine.train_step", "results": res,
            "roofline": {"bound": "tensor (tf32)", "peak_bf16_tflops": peaks["tensor_tflops"],
                         "note": "tf32 tensor rate is half the 16-bit rate; the 256-row reference batch is launch-latency "
                                 "bound (about 45 launches per step)"},
            "cpu_baseline": cpu}
    out_fd.write(json.dumps(line) + "\n")
    out_fd.flush()


if __name__ == "__main__":
    main()
