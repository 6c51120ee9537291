"""Writes tests/golden/sampler_golden.npz by running the REFERENCE's own rollout loop end to end on the CPU:
milo/milo/sampler.py::get_samples driving the reference SimEnv (gym-simenv/gym_simenv/envs/sim_env.py), a reference
DynamicsEnsemble (milo/milo/dynamics.py) and the reference Gaussian MLP policy (mjrl/mjrl/policies/gaussian_mlp.py).

Run in the build container only (needs /root/reference):  python tests/golden/make_sampler_golden.py

Stubs: exactly those of make_simenv_golden.py (gym base classes, the SWIG simulator) plus tkinter for dynamics.py.
The exploration noise is reproducible from the stored seeds: get_samples calls np.random.seed(seed + k) before
trajectory k (sampler.py:38) and MLP.get_action then draws uniform() followed by randn(action_dim) per step
(gaussian_mlp.py:95-104); resets draw their clip time from the env's own RandomState (sim_env.py:276).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import make_golden as mg  # noqa: E402
import make_simenv_golden as msg  # noqa: E402
from oracle import imitation_oracle as io  # noqa: E402
from tests import helpers as H  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler_golden.npz")
S, A = msg.S, msg.A


def main():
    ref = mg.load_reference()
    SimEnv = msg.load_reference_simenv()
    sys.path.insert(0, os.path.join(mg.REF, "mjrl"))
    from mjrl.policies.gaussian_mlp import MLP
    spec = importlib.util.spec_from_file_location("ref_sampler", os.path.join(mg.REF, "milo", "milo", "sampler.py"))
    sampler = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sampler)
    msg.FakeSimulator.clip = io.Clip(H.spinkick_raw(), io.HUMANOID3D, "wrap")
    os.chdir(msg.DM_ROOT)

    N, hidden, horizon, seed, n_traj = 3, [32, 32], 6, 40, 6
    s, a, s2 = mg.synth_dataset(512, S, A, 0)
    ens = mg.build_ensemble(ref, S, A, N, hidden, True, "relu", ref["datasets"].AmpDataset(s, a, s2))
    env = SimEnv(ens, deepmimic_args=msg.ARG_FILE, horizon=horizon, seed=1)
    pol = MLP(S, A, hidden_sizes=(32, 32), seed=123, init_log_std=-1.0, min_log_std=-2.5)

    paths, n_samples = sampler.get_samples(env, pol, n_traj, seed, mode="trajectories", eval_mode=False)
    assert len(paths) == n_traj and n_samples == sum(len(p["rewards"]) for p in paths)
    T = max(len(p["rewards"]) for p in paths)
    obs, nxt = np.zeros((n_traj, T, S)), np.zeros((n_traj, T, S))
    act, mean = np.zeros((n_traj, T, A)), np.zeros((n_traj, T, A))
    length = np.zeros(n_traj, dtype=np.int64)
    for k, p in enumerate(paths):
        n = len(p["rewards"])
        length[k] = n
        obs[k, :n], nxt[k, :n], act[k, :n] = p["observations"], p["next_observations"], p["actions"]
        mean[k, :n] = p["agent_infos"]["mean"]
        assert p["terminated"] and np.all(p["rewards"] == 0)
        assert np.array_equal(p["observations"][1:], p["next_observations"][:-1])
    out = dict(N=np.int64(N), hidden=np.array(hidden), horizon=np.int64(horizon), seed=np.int64(seed),
               dataset_seed=np.int64(0), dataset_rows=np.int64(512), base_seed=np.int64(100),
               observations=obs, next_observations=nxt, actions=act, means=mean, length=length,
               member=np.array([(k + 1) % N for k in range(n_traj)]),      # sim_env.py:282-283, checked below
               log_std=pol.log_std.data.numpy().copy(),
               weight_checksum=np.array([[float(l.weight.detach().double().abs().sum()) for l in m.model.fc_layers]
                                         for m in ens.models]))
    assert env.reset_counter == n_traj % N
    for i, l in enumerate(pol.model.fc_layers):
        out[f"pol_w{i}"], out[f"pol_b{i}"] = l.weight.data.numpy().copy(), l.bias.data.numpy().copy()
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes; path lengths", length, "samples", n_samples)


if __name__ == "__main__":
    main()
