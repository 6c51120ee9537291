"""the batched on-device rollout loop (policy forward + env step + MILO cost +
auto-reset over a horizon), measured like bench.py measures the step.

    python tools/bench_rollout.py [--envs 40000] [--horizon 32] [--iters 5] [--graph]

Prints one JSON line: env-steps/s through amp_extensions_b200.rollout.DeviceRollout.collect (device timed), the
same with the loop replayed from a CUDA graph, a small-batch pair (E = 1024) where launch overhead dominates, and
the oracle's batched restatement of the reference sampler on the host cores on a bounded sample.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


class _FC:
    def __init__(self, sizes, seed):
        g = torch.Generator().manual_seed(seed)
        self.fc_layers = [torch.nn.Linear(sizes[i], sizes[i + 1]) for i in range(len(sizes) - 1)]
        for l in self.fc_layers:
            l.weight.data = torch.randn(l.weight.shape, generator=g) * (1.0 / l.weight.shape[1]) ** 0.5
            l.bias.data = torch.zeros(l.bias.shape)
        self.fc_layers[-1].weight.data *= 1e-2   # gaussian_mlp.py:38-39
        self.nonlinearity = torch.tanh
        self.in_shift, self.in_scale = torch.zeros(sizes[0]), torch.ones(sizes[0])
        self.out_shift, self.out_scale = torch.zeros(sizes[-1]), torch.ones(sizes[-1])


class _Policy:
    def __init__(self, obs, act, seed=123):
        self.model = _FC((obs, 32, 32, act), seed)      # run.py's policy_size = (32, 32)
        self.log_std = torch.full((act,), -0.5)


def timed(fn, iters, dev):
    fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=40000)
    ap.add_argument("--horizon", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    out_fd = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import bench as B
    from amp_extensions_b200 import AmpDataset, DynamicsEnsemble, RBFLinearCost, VecSimEnv, _lib
    from amp_extensions_b200.rollout import DeviceRollout
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    S, A = B.S_DIM, B.A_DIM
    ds = AmpDataset(*B.synth_dataset(8192, 0))
    ens = DynamicsEnsemble(S, A, ds, None, num_models=B.N_MODELS, hidden_sizes=B.HIDDEN, dense_connect=True,
                           transform=True, base_seed=100, device=dev)
    ens.train_dataset = AmpDataset(ds.states[:1024], ds.actions[:1024], ds.next_states[:1024])
    ens.compute_threshold()
    cost = RBFLinearCost(B.synth_expert(4096, 2), feature_dim=B.RFF_DIM, input_type="ss", bw_quantile=B.BW_QUANTILE,
                         lambda_b=B.LAMBDA_B, seed=100, device=dev)
    cost.fit_cost(torch.cat([ds.states[:1024], ds.next_states[:1024]], dim=1))
    pool = ds.states[:4096] * 0.3
    pool[:, 0] = 1.5
    pol = _Policy(S, A)
    res = {}
    for E in (args.envs, 1024):
        env = VecSimEnv(ens, E, horizon=300, reset_states=pool, seed=1, cost=cost)
        env.reset()
        ro = DeviceRollout(env, pol, seed=0)
        T = args.horizon
        l0 = _lib.launch_count()
        ms = timed(lambda: ro.collect(T), args.iters, dev)
        launches = (_lib.launch_count() - l0) // (args.iters + 1)
        entry = {"envs": E, "horizon": T, "ms_per_rollout": ms, "env_steps_per_s": E * T / (ms * 1e-3),
                 "launches_per_rollout": launches}
        try:
            msg = timed(lambda: ro.collect(T, graph=True), args.iters, dev)
            entry["graph"] = {"ms_per_rollout": msg, "env_steps_per_s": E * T / (msg * 1e-3)}
        except Exception as e:  # report, do not hide
            entry["graph"] = {"error": repr(e)[:300]}
        res[f"E{E}"] = entry
        del ro, env
    cpu = None
    if not args.skip_cpu:
        from oracle import milo_oracle as mo
        from oracle import rollout_oracle as rlo
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        Ec, Tc = 1024, 4
        ws, bs = mo.init_ensemble(S, A, B.HIDDEN, B.N_MODELS, dense_connect=True, base_seed=100)
        tf = mo.get_transformations(ds.states, ds.actions, ds.next_states)
        oc = mo.RffCostOracle(B.synth_expert(4096, 2), feature_dim=B.RFF_DIM, input_type="ss", bw_quantile=B.BW_QUANTILE,
                              lambda_b=B.LAMBDA_B, seed=100)
        oc.w = cost.w
        pd = dict(ws=[l.weight.data for l in pol.model.fc_layers], bs=[l.bias.data for l in pol.model.fc_layers],
                  log_std=pol.log_std.numpy())
        rng = np.random.default_rng(0)
        noise = rng.standard_normal((Tc, Ec, A)).astype(np.float32)
        pick = rng.integers(0, 4096, (Tc, Ec)).astype(np.int32)
        st = pool[:Ec].numpy()
        args_o = (ws, bs, tf, pd, st, np.zeros(Ec, np.int64), np.zeros(Ec, np.int64), pool.numpy(), noise, pick)
        rlo.rollout(*args_o, cost=oc, threshold=ens.threshold)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 10.0:
            rlo.rollout(*args_o, cost=oc, threshold=ens.threshold)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": Ec * Tc / dt, "unit": "env-steps/s", "cores": threads, "kind": "port",
               "sample": f"{reps} x ({Ec} envs x {Tc} steps), oracle restatement of the reference sampler loop, batched"}
    line = {"workload": "on-device rollout: Gaussian MLP policy (226-32-32-28) + env step + IPM/RFF cost + auto-reset",
            "api": "amp_extensions_b200.rollout.DeviceRollout.collect", "results": res, "cpu_baseline": cpu}
    out_fd.write(json.dumps(line) + "\n")
    out_fd.flush()


if __name__ == "__main__":
    main()
