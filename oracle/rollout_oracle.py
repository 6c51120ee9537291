"""CPU oracle for the rollout loop around the env step — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in torch-CPU fp32 / numpy float64 (the reference's own arithmetic), the functions either side of the
env step that SURVEY.md section 8(f) ranks next: the Gaussian MLP policy forward, the sampler loop, and the
discounted-sum / GAE post-processing.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
this module.

Pinning: tests/golden/rollout_golden.npz was produced by tests/golden/make_rollout_golden.py, which imports the
reference's own mjrl/mjrl/utils/{fc_network,process_samples}.py and mjrl/mjrl/policies/gaussian_mlp.py in the
build container and stores their outputs; tests/test_rollout_oracle.py checks the functions below against those
vectors.  The sampler loop itself (milo/milo/sampler.py::get_samples) is pinned as well:
tests/golden/make_sampler_golden.py runs it unmodified over the reference SimEnv (gym / simulator imports stubbed as
in make_simenv_golden.py), a reference DynamicsEnsemble and the reference MLP policy; `rollout` replays those
trajectories from the stored seeds' exploration draws (tests/test_rollout_oracle.py, tests/golden/sampler_golden.npz).

Reference files (paths under the reference tree):
  FC  = mjrl/mjrl/utils/fc_network.py          GM = mjrl/mjrl/policies/gaussian_mlp.py
  PS  = mjrl/mjrl/utils/process_samples.py     SMP = milo/milo/sampler.py
  SE  = gym-simenv/gym_simenv/envs/sim_env.py  BR = mjrl/mjrl/algos/batch_reinforce.py
"""
import numpy as np
import torch

from . import milo_oracle as mo


def fc_forward(ws, bs, x, nonlinearity="tanh", in_shift=None, in_scale=None, out_shift=None, out_scale=None):
    """FCNetwork.forward (FC:42-55): (x - in_shift) / (in_scale + 1e-8), hidden layers with the nonlinearity,
    last layer linear, * out_scale + out_shift.  fp32 torch, like the reference."""
    x = torch.as_tensor(x, dtype=torch.float32)
    obs_dim, act_dim = ws[0].shape[1], ws[-1].shape[0]
    in_shift = torch.zeros(obs_dim) if in_shift is None else torch.as_tensor(in_shift, dtype=torch.float32)
    in_scale = torch.ones(obs_dim) if in_scale is None else torch.as_tensor(in_scale, dtype=torch.float32)
    out_shift = torch.zeros(act_dim) if out_shift is None else torch.as_tensor(out_shift, dtype=torch.float32)
    out_scale = torch.ones(act_dim) if out_scale is None else torch.as_tensor(out_scale, dtype=torch.float32)
    act = torch.relu if nonlinearity == "relu" else torch.tanh
    out = (x - in_shift) / (in_scale + 1e-8)
    for i in range(len(ws) - 1):
        out = act(torch.nn.functional.linear(out, ws[i], bs[i]))
    out = torch.nn.functional.linear(out, ws[-1], bs[-1])
    return out * out_scale + out_shift


def get_action(mean, log_std, noise):
    """MLP.get_action (GM:95-104) with the standard-normal draw made explicit: float64
    `mean + exp(log_std_val) * noise`, log_std_val = float64(log_std) (GM:53)."""
    mean = np.asarray(mean)
    return mean + np.exp(np.float64(np.asarray(log_std))) * np.asarray(noise, dtype=np.float64)


def discount_sum(x, gamma, terminal=0.0):
    """PS:37-45."""
    y = []
    run_sum = terminal
    for t in range(len(x) - 1, -1, -1):
        run_sum = x[t] + gamma * run_sum
        y.append(run_sum)
    return np.array(y[::-1])


def gae_advantages(rewards, baseline, terminated, gamma, gae_lambda):
    """compute_advantages, GAE branch, 1-D baseline (PS:21-30)."""
    b = np.asarray(baseline)
    b1 = np.append(b, 0.0 if terminated else b[-1])
    td = np.asarray(rewards) + gamma * b1[1:] - b1[:-1]
    return discount_sum(td, gamma * gae_lambda)


def rollout(ws, bs, tfs, policy, env_state, member, num_steps, pool, noise, pick, horizon=300, cost=None,
            threshold=1.0, n_models=None):
    """The batched form of SMP:36-66 on top of the pinned pieces, for parity with DeviceRollout.collect:
    for t < T: a = get_action(policy(o)), (o', done) = SimEnv.step (SE:140-173 through milo_oracle), optional
    MILO cost (LC:111-152), then envs that are done reset from pool[pick] with num_steps = 0 and the member index
    advanced (SE:270-285).

    policy: dict(ws, bs, nonlinearity, in_shift, in_scale, out_shift, out_scale, log_std)
    cost:   milo_oracle.RffCostOracle with .w set, or None.
    Returns dict of time-major numpy arrays.
    """
    T, E = noise.shape[0] if noise is not None else pick.shape[0], env_state.shape[0]
    n_models = n_models or len(ws)
    o = np.asarray(env_state, dtype=np.float32).copy()
    member = np.asarray(member).copy()
    steps = np.asarray(num_steps).astype(np.int64).copy()
    out = dict(observations=[], next_observations=[], actions=[], means=[], disc=[], done=[], cost=[], ipm=[],
               bonus=[])
    for t in range(T):
        mean = fc_forward(policy["ws"], policy["bs"], torch.from_numpy(o), policy.get("nonlinearity", "tanh"),
                          policy.get("in_shift"), policy.get("in_scale"), policy.get("out_shift"),
                          policy.get("out_scale")).numpy()
        a = mean if noise is None else get_action(mean, policy["log_std"], noise[t]).astype(np.float32)
        ot, at = torch.from_numpy(o), torch.from_numpy(np.asarray(a, dtype=np.float32))
        preds = mo.ensemble_forward(ws, bs, tfs, ot, at)
        active = preds[torch.from_numpy(member).long(), torch.arange(E)]
        nxt, steps, done = mo.simenv_step(o.astype(np.float64), active.numpy(), steps, horizon=horizon)
        nxt32 = nxt.astype(np.float32)
        disc = mo.discrepancy_from_preds(preds)
        out["observations"].append(o.copy())
        out["next_observations"].append(nxt32)
        out["actions"].append(np.asarray(a, dtype=np.float32))
        out["means"].append(mean)
        out["disc"].append(disc.numpy())
        out["done"].append(done.copy())
        if cost is not None:
            c, info = cost.get_bonus_costs(ot, at, disc, threshold, next_states=torch.from_numpy(nxt32))
            out["cost"].append(c[:, 0].numpy())
            out["ipm"].append(info["ipm"][:, 0].numpy())
            out["bonus"].append(info["bonus"][:, 0].numpy())
        o = nxt32.copy()
        idx = np.flatnonzero(done)
        if idx.size:
            o[idx] = np.asarray(pool, dtype=np.float32)[np.asarray(pick[t])[idx] % len(pool)]
            steps[idx] = 0
            member[idx] = (member[idx] + 1) % n_models
    res = {k: np.stack(v) for k, v in out.items() if v}
    res["final_state"], res["member"], res["num_steps"] = o, member, steps
    return res
