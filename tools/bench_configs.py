"""BASELINE.json configs[3] and configs[4] (the bench.py line is configs[1]; configs[2] is tools/bench_imitation.py).

    python tools/bench_configs.py --config 4 [--envs-total 1048576]     # fused step + cost + imitation reward
    python tools/bench_configs.py --config 5 [--envs-per-gpu 524288]    # 8 x (1024 x 4) ensemble scale sweep
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_configs.py --config 4   # sharded over N GPUs

config 4: "fused ensemble step + discrepancy quantile + MILO cost + imitation reward, 1M parallel envs sharded over
2/4/8 B200": the env range [0, envs_total) is sharded by index; every step each rank runs simstep_step_cost on its
envs and simstep_imitation_reward on as many synthetic humanoid3d poses; once per `--quantile-every` steps the
global bw_quantile-style discrepancy quantile is taken over ALL envs with histogram all-reduces (parallel.py).
config 5: "8-model ensemble (hidden 1024 x 4) scale sweep, 4M envs at 8 x B200 with global quantile via NCCL":
weak-scaled, 512 Ki envs per GPU by default.

Prints one JSON line on rank 0: env-steps/s (device timed, max over ranks), per-kernel device times, ensemble GEMM
TFLOP/s against the measured tensor peak, imitation GB/s against the measured HBM peak, quantile latency.
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
sys.path.insert(0, os.path.join(ROOT, "tools"))

S, A = 226, 28


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, choices=[4, 5], required=True)
    ap.add_argument("--envs-total", type=int, default=1 << 20)
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 19)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--quantile-every", type=int, default=10)
    ap.add_argument("--precision", default="fp16")
    args = ap.parse_args()
    out_fd = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import bench as B
    from bench_imitation import synth_poses
    from amp_extensions_b200 import AmpDataset, DynamicsEnsemble, ImitationReward, RBFLinearCost, parallel
    from amp_extensions_b200.engine import HumanoidTermination

    if args.config == 4:
        n_models, hidden = 4, [512] * 4
        lo, hi = parallel.shard_range(args.envs_total)
        E = hi - lo
    else:
        n_models, hidden = 8, [1024] * 4
        E = args.envs_per_gpu
    ds = AmpDataset(*B.synth_dataset(8192, 0))
    ens = DynamicsEnsemble(S, A, ds, None, num_models=n_models, hidden_sizes=hidden, dense_connect=True,
                           transform=True, base_seed=100, precision=args.precision, device=dev)
    eng = ens.engine()
    eng.set_termination(HumanoidTermination(horizon=300))
    cost = RBFLinearCost(B.synth_expert(4096, 2), feature_dim=512, input_type="ss", bw_quantile=0.1, lambda_b=0.0025,
                         seed=100, precision=args.precision, device=dev)
    eng.load_rff(cost.rff.weight.data, cost.rff.bias.data, split=True)
    ens.train_dataset = AmpDataset(ds.states[rank::world][:1024], ds.actions[rank::world][:1024],
                                   ds.next_states[rank::world][:1024])
    threshold = parallel.global_threshold(ens)

    g = torch.Generator(device=dev).manual_seed(1 + rank)
    ring = 2
    states = [torch.randn(E, S, device=dev, generator=g) for _ in range(ring)]
    actions = [torch.randn(E, A, device=dev, generator=g) for _ in range(ring)]
    member = torch.randint(0, n_models, (E,), device=dev, generator=g, dtype=torch.int32)
    steps = torch.zeros(E, device=dev, dtype=torch.int32)
    nxt = torch.empty(E, S, device=dev)
    disc, cst, ipm, bonus = (torch.empty(E, device=dev) for _ in range(4))
    done = torch.empty(E, device=dev, dtype=torch.uint8)
    eng.step(states[0][:4096], actions[0][:4096], member[:4096], steps[:4096].clone(), next_state=nxt[:4096],
             disc=disc[:4096], done=done[:4096])
    w = parallel.global_fit_cost(cost, torch.cat([states[0][:1024], nxt[:1024]], dim=1)).to(dev)

    imit = None
    if args.config == 4:
        imit = ImitationReward(device=dev)
        poses = [synth_poses(imit, E, dev, seed=3 + 7 * rank + i) for i in range(ring)]

    def one_step(i):
        k = i % ring
        eng.step_cost(states[k], actions[k], member, steps, w, 0.0025, threshold, next_state=nxt, disc=disc, done=done,
                      cost=cst, ipm=ipm, bonus=bonus)
        if imit is not None:
            p, v, t = poses[k]
            return imit.reward(p, v, t)
        return None

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        one_step(i)
    q = parallel.global_quantile(disc, 0.9, engine=eng)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_q, n_q = 0.0, 0
    e0.record()
    t_wall = time.perf_counter()
    for i in range(args.steps):
        one_step(args.warmup + i)
        if (i + 1) % args.quantile_every == 0:
            tq0 = time.perf_counter()
            q = parallel.global_quantile(disc, 0.9, engine=eng)   # host-synchronous: inside the timed region on purpose
            t_q += time.perf_counter() - tq0
            n_q += 1
    e1.record()
    sync()
    wall = time.perf_counter() - t_wall
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(E)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms = float(ms.item())
    value = float(tot.item()) * args.steps / (total_ms * 1e-3)

    prof_steps = min(args.steps, 10)
    eng.profile_enable(True)
    for i in range(prof_steps):
        one_step(i)
    prof = eng.profile_read(reset=True)
    eng.profile_enable(False)
    per_step = {k: v[0] / prof_steps for k, v in prof.items() if v[1] > 0}
    imit_ms = None
    if imit is not None:
        torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(prof_steps):
            p, v, t = poses[i % ring]
            imit.reward(p, v, t)
        a1.record()
        torch.cuda.synchronize(dev)
        imit_ms = a0.elapsed_time(a1) / prof_steps

    if rank == 0:
        peaks = B.measured_peaks()
        flop = B.flops_per_env_step(n_models, hidden)
        gemm_ms = per_step.get("ensemble_gemm")
        line = {
            "config": args.config, "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world,
            "steps": args.steps, "ms_per_step": total_ms / args.steps, "wall_ms_per_step": wall / args.steps * 1e3,
            "envs_per_gpu": E, "envs_total": int(tot.item()),
            "ensemble": f"{n_models} x ({hidden[0]} x {len(hidden)}) dense_connect, S=226 A=28",
            "with": ["ensemble step", "discrepancy", "termination", "IPM/RFF cost + bonus"] +
                    (["imitation reward (1 pose per env)"] if imit is not None else []) +
                    [f"global 0.9-quantile of the discrepancy over all ranks every {args.quantile_every} steps"],
            "quantile": {"value": q, "ms_per_call": (t_q / n_q * 1e3) if n_q else None, "calls": n_q},
            "threshold": threshold, "kernels_ms_per_step": per_step,
            "roofline": {"bound": "tensor", "achieved": flop * E / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                         "peak": peaks["tensor_tflops"], "unit": "TFLOP/s", "algorithmic_flop_per_env_step": flop},
            "dtype": args.precision,
        }
        if line["roofline"]["achieved"]:
            line["roofline"]["frac"] = line["roofline"]["achieved"] / peaks["tensor_tflops"]
        if imit_ms:
            gbs = 352 * E / (imit_ms * 1e-3) / 1e9
            line["imitation"] = {"ms_per_step": imit_ms, "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": gbs / peaks["hbm_gbs"]}
        out_fd.write(json.dumps(line) + "\n")
        out_fd.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
