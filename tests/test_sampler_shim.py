"""amp_extensions_b200.sampler (the drop-in for milo/milo/sampler.py) against the reference's own sampler run.

tests/golden/sampler_golden.npz holds what the reference's get_samples produced over the reference SimEnv,
DynamicsEnsemble and Gaussian MLP policy (tests/golden/make_sampler_golden.py).  Here the shim's HOST logic - per
trajectory seeding of env and numpy, the uniform()/randn() draw order of MLP.get_action, the member round-robin,
cutting every env column at its first `done`, the quota rule, the path dict layout - runs on the CPU with a backend
built on the oracle's batched rollout, and must reproduce those trajectories.  The GPU backend is covered by
tests/test_rollout_gpu.py::test_sampler_shim_matches_the_reference_sampler_on_gpu.
"""
import numpy as np
import torch

from amp_extensions_b200 import sampler as shim
from oracle import milo_oracle as mo
from oracle import rollout_oracle as ro
from tests import helpers as H
from tests.test_rollout_oracle import sampler_golden


class GoldenEnv:
    """The surface the sampler touches, with reset() handing out the golden run's initial states: the reference
    drew them from its simulator stand-in with the env generator seeded by seed + k (sim_env.py:122-132, 276)."""

    def __init__(self, g):
        self.g = g
        self.horizon = int(g["horizon"])
        self.state_size, self.action_size = 226, 28
        self.reset_counter = 0
        self.dynamic_ensemble = type("E", (), {"models": [None] * int(g["N"])})()
        self.seeds = []

    def seed_env(self, seed):
        self.seeds.append(seed)
        self._k = seed - int(self.g["seed"])   # trajectory counter, 1-based

    def reset(self):
        self.reset_counter = (self.reset_counter + 1) % len(self.dynamic_ensemble.models)
        k = min(self._k, self.g["observations"].shape[0]) - 1
        return self.g["observations"][k, 0].copy()


class OracleBackend:
    def __init__(self, g):
        N, hidden = int(g["N"]), [int(h) for h in g["hidden"]]
        s, a, s2 = H.synth_dataset(int(g["dataset_rows"]), 226, 28, int(g["dataset_seed"]))
        self.ws, self.bs = mo.init_ensemble(226, 28, hidden, N, base_seed=int(g["base_seed"]), dense_connect=True)
        self.tf = mo.get_transformations(s, a, s2)
        self.policy = dict(ws=[torch.from_numpy(g[f"pol_w{i}"]) for i in range(3)],
                           bs=[torch.from_numpy(g[f"pol_b{i}"]) for i in range(3)], log_std=g["log_std"])
        self.N, self.horizon = N, int(g["horizon"])
        self.calls = 0

    def rollout(self, ob0, member, noise, T):
        self.calls += 1
        E = ob0.shape[0]
        res = ro.rollout(self.ws, self.bs, self.tf, self.policy, ob0, member, np.zeros(E, dtype=np.int64), ob0, noise,
                         np.zeros((T, E), dtype=np.int64), horizon=self.horizon, n_models=self.N)
        res["log_std"] = np.asarray(self.policy["log_std"], dtype=np.float32)
        return res


def check_paths(paths, g, first=0):
    for j, p in enumerate(paths):
        k = first + j
        n = int(g["length"][k])
        assert len(p["rewards"]) == n and p["terminated"] and np.all(p["rewards"] == 0), (k, len(p["rewards"]), n)
        for name in ("observations", "next_observations", "actions"):
            assert p[name].dtype == np.float64 and p[name].shape[0] == n
            np.testing.assert_allclose(p[name], g[name][k, :n], atol=2e-5)
        np.testing.assert_allclose(p["agent_infos"]["mean"], g["means"][k, :n], atol=2e-5)
        np.testing.assert_allclose(p["agent_infos"]["evaluation"], g["means"][k, :n], atol=2e-5)
        assert p["agent_infos"]["log_std"].shape == (n, 28)
        assert len(p["env_infos"]) == n and all(i["valid"] for i in p["env_infos"])


def test_get_samples_reproduces_the_reference_run_in_trajectory_mode():
    g = sampler_golden()
    env, be = GoldenEnv(g), OracleBackend(g)
    n_traj = g["actions"].shape[0]
    paths, n = shim.get_samples(env, None, n_traj, int(g["seed"]), mode="trajectories", backend=be)
    assert len(paths) == n_traj and n == int(g["length"].sum())
    assert env.seeds == [int(g["seed"]) + k for k in range(1, n_traj + 1)]      # sampler.py:36-38
    check_paths(paths, g)
    assert be.calls == 1                                                         # one batched rollout, not six
    assert env.reset_counter == n_traj % int(g["N"])                             # as if the env had been reset 6 times


def test_get_samples_quota_rule_in_sample_mode():
    """sampler.py:30-34, 79-82: whole trajectories until the sample count reaches the quota."""
    g = sampler_golden()
    lengths = [int(x) for x in g["length"]]                                      # 2 1 6 1 1 6
    for quota in (1, 2, 3, 4, 9, 10, 11):
        env, be = GoldenEnv(g), OracleBackend(g)
        paths, n = shim.get_samples(env, None, quota, int(g["seed"]), mode="samples", backend=be)
        want, tot = 0, 0
        while tot < quota:
            tot += lengths[want]
            want += 1
        assert len(paths) == want and n == tot, (quota, len(paths), want, n, tot)
        check_paths(paths, g)


def test_sample_points_orders_workers_and_restarts_the_member_counter():
    """sampler.py:111-128: worker i gets seed 12345 + base_seed * i and its own env copy (reset_counter 0); results
    are concatenated worker by worker.  With base_seed chosen so that worker 0's seed equals the golden run's, worker
    0's paths are the golden trajectories; eval_mode returns the policy mean as the action."""
    g = sampler_golden()
    seed = int(g["seed"])

    class Env(GoldenEnv):
        def seed_env(self, s):
            self.seeds.append(s)
            self._k = s - 12345 if s - 12345 <= 6 else 1

    env, be = Env(g), OracleBackend(g)
    g2 = dict(g)
    g2["seed"] = np.int64(12345)
    # worker seeds: 12345, 12345 + base_seed (far away: its trajectories start from golden state 0 again)
    paths = shim.sample_points(env, None, 4, base_seed=1000, num_workers=2, mode="trajectories", eval_mode=True,
                               backend=be)
    assert len(paths) == 4
    assert env.seeds[:2] == [12346, 12347] and env.seeds[2:4] == [13346, 13347]
    for p in paths:
        np.testing.assert_array_equal(p["actions"], p["agent_infos"]["mean"].astype(np.float64))
    assert env.reset_counter == 0                                                # the caller's env is left alone
