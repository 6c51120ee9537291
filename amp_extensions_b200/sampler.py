"""Drop-in for milo/milo/sampler.py (reference lines 8-130): `get_samples` / `sample_points` with the reference's
signatures and return values, evaluated as ONE batched rollout on the device instead of one `env.step` at a time in
a multiprocessing Pool.

What is kept from the reference, draw for draw:
  * trajectory k of a worker is seeded with `seed + k` for both the env (`env.seed_env`, sampler.py:37) and numpy's
    global generator (sampler.py:38); the env's reset draws come from the env's own generator (sim_env.py:276), the
    policy's exploration draws from numpy's: `uniform()` then `randn(m)` per step (gaussian_mlp.py:95-104);
  * `sample_points` gives worker i the seed `12345 + base_seed * i` and `ceil(num_to_collect / num_workers)` samples
    or trajectories (sampler.py:111-115); a worker's env copy starts from the pickled constructor arguments
    (EzPickle, sim_env.py:48), i.e. with reset_counter 0, so its first trajectory runs on member 1;
  * a worker stops after the first trajectory that brings it to its quota and keeps every complete trajectory up to
    there (sampler.py:30-34, 79-82); paths are dicts of float64 numpy arrays in the layout batch_reinforce.py reads.

What changes: all trajectories of all workers advance together, one env column each; every step is three library
calls on device buffers (rollout.DeviceRollout), and nothing crosses PCIe until the paths are assembled.  The
trajectories a worker would not have started (its quota was already met) are computed speculatively and dropped.
A policy with eps > 0 (uniform random actions, gaussian_mlp.py:98-99) falls back to the sequential loop.

`backend` is the thing that advances a batch of trajectories; tests inject a CPU one built on the oracle.
"""
import copy
import math
import time

import numpy as np
import torch


class DeviceBackend:
    """Advances E trajectories for T steps on the GPU through rollout.DeviceRollout."""

    def __init__(self, ensemble, policy, termination=None, horizon=300, enable_velocity_check=False):
        self.ensemble, self.policy = ensemble, policy
        self.termination, self.horizon, self.vel = termination, horizon, enable_velocity_check
        self._env, self._ro = None, None

    def rollout(self, ob0, member, noise, T):
        from .rollout import DeviceRollout
        from .sim_env import VecSimEnv
        E = ob0.shape[0]
        ob0 = torch.as_tensor(ob0, dtype=torch.float32)
        if self._env is None or self._env.num_envs != E:
            self._env = VecSimEnv(self.ensemble, E, termination=self.termination, horizon=self.horizon,
                                  enable_velocity_check=self.vel, reset_states=ob0)
            self._ro = DeviceRollout(self._env, self.policy, seed=0)
        env, ro = self._env, self._ro
        ro.pool = ob0.to(env.device).contiguous()
        env.ob.copy_(ro.pool)
        env.num_steps.zero_()
        env.member.copy_(torch.as_tensor(member, dtype=torch.int32))
        dev = env.device
        nz = None if noise is None else torch.as_tensor(noise, dtype=torch.float32).to(dev)
        b = ro.collect(T, eval_mode=noise is None, noise=nz, pick=torch.zeros((T, E), dtype=torch.int32, device=dev))
        out = {k: getattr(b, k).cpu().numpy() for k in ("observations", "next_observations", "actions", "means", "done",
                                                        "disc")}
        out["log_std"] = np.asarray(ro.log_std, dtype=np.float32).ravel()
        return out


def _first_done(done_col, horizon):
    idx = np.flatnonzero(done_col)
    return int(idx[0]) + 1 if idx.size else min(len(done_col), horizon)


def _draws(seed, T, m):
    """The numpy draws of one trajectory: np.random.seed(seed), then per step uniform() and randn(m)."""
    rs = np.random.RandomState(seed)
    noise = np.zeros((T, m))
    for t in range(T):
        rs.uniform()
        noise[t] = rs.randn(m)
    return noise


def _collect(env, policy, quotas, seeds, mode, eval_mode, backend, start_counter=0):
    """quotas[i], seeds[i]: worker i's share and base seed; start_counter: resets the worker envs have seen before.
    Returns (paths per worker, samples per worker)."""
    horizon = int(env.horizon)
    m = int(env.action_size)
    n_workers = len(quotas)
    n_models = len(env.dynamic_ensemble.models)
    if backend is None:
        backend = DeviceBackend(env.dynamic_ensemble, policy, termination=getattr(env, "_termination", None),
                                horizon=horizon, enable_velocity_check=getattr(env, "enable_velocity_check", False))
    results = [[] for _ in range(n_workers)]
    counts = [0] * n_workers          # samples (or trajectories) each worker has so far
    next_k = [1] * n_workers          # the worker's seed counter (sampler.py:35)
    while True:
        # how many more trajectories each worker needs at least; a wave runs that many per worker
        lanes = []
        for i in range(n_workers):
            left = quotas[i] - counts[i]
            if left <= 0:
                continue
            want = left if mode == "trajectories" else max(1, math.ceil(left / horizon))
            for _ in range(want):
                lanes.append((i, next_k[i]))
                next_k[i] += 1
        if not lanes:
            break
        # initial states come from the env's own reset(), seeded per trajectory exactly as the reference does
        ob0 = np.zeros((len(lanes), env.state_size))
        member = np.zeros(len(lanes), dtype=np.int64)
        for j, (i, k) in enumerate(lanes):
            env.seed_env(seeds[i] + k)
            env.reset_counter = (start_counter + k - 1) % n_models  # resets this worker's env has seen before
            ob0[j] = env.reset()
            member[j] = env.reset_counter
        noise = None
        if not eval_mode:
            noise = np.stack([_draws(seeds[i] + k, horizon, m) for (i, k) in lanes], axis=1)  # [T, E, m]
        out = backend.rollout(ob0, member, noise, horizon)
        log_std = np.asarray(out.get("log_std", getattr(policy, "log_std_val", np.zeros(m))), dtype=np.float64).ravel()
        for j, (i, k) in enumerate(lanes):
            if counts[i] >= quotas[i]:
                continue  # speculative trajectory: the worker had already stopped
            n = _first_done(out["done"][:, j], horizon)
            mean = out["means"][:n, j]
            act = (mean if eval_mode else out["actions"][:n, j]).astype(np.float64)
            infos = [{"valid": True, "disc": float(out["disc"][t, j])} for t in range(n)] if "disc" in out else \
                    [{"valid": True} for _ in range(n)]
            results[i].append(dict(
                observations=out["observations"][:n, j].astype(np.float64),
                next_observations=out["next_observations"][:n, j].astype(np.float64),
                actions=act, rewards=np.zeros(n),
                agent_infos=dict(mean=mean, log_std=np.tile(log_std, (n, 1)), evaluation=mean),
                env_infos=infos, terminated=True))
            counts[i] += 1 if mode == "trajectories" else n
    samples = [sum(len(p["rewards"]) for p in r) for r in results]
    return results, samples


def _sequential(env, policy, num_to_collect, seed, mode, eval_mode, deepmimic):
    """The reference loop itself (sampler.py:26-84), for policies the batched path does not cover."""
    paths, n_paths, n_samples, k = [], 0, 0, 0
    while (n_paths if mode == "trajectories" else n_samples) < num_to_collect:
        k += 1
        env.seed_env(seed + k)
        np.random.seed(seed + k)
        obs, acts, rews, ainfos, nobs, einfos = [], [], [], [], [], []
        o, done, valid = env.reset(), False, True
        while not done:
            a, info = policy.get_action(o)
            a = info["evaluation"] if eval_mode else a
            no, r, done, einfo = env.step(a)
            obs.append(o); nobs.append(no); acts.append(a); rews.append(r); ainfos.append(info); einfos.append(einfo)
            if deepmimic and not einfo["valid"]:
                done, valid = True, False
            o = no
        if valid:
            paths.append(dict(observations=np.array(obs), next_observations=np.array(nobs), actions=np.array(acts),
                              rewards=np.array(rews),
                              agent_infos={key: np.array([x[key] for x in ainfos]) for key in ainfos[0]},
                              env_infos=einfos, terminated=done))
            n_paths += 1
            n_samples += len(obs)
    return paths, n_samples


def get_samples(env, policy, num_to_collect, seed, mode="samples", eval_mode=False, deepmimic=False, backend=None):
    """sampler.py:8-84 for one worker: returns (paths, samples_collected)."""
    assert mode in ("samples", "trajectories")
    if getattr(policy, "eps", 0.0):
        return _sequential(env, policy, num_to_collect, seed, mode, eval_mode, deepmimic)
    saved = env.reset_counter
    results, samples = _collect(env, policy, [num_to_collect], [seed], mode, eval_mode, backend, start_counter=saved)
    env.reset_counter = (saved + len(results[0])) % len(env.dynamic_ensemble.models)
    return results[0], samples[0]


def sample_points(env, policy, num_to_collect, base_seed, num_workers=4, mode="samples", eval_mode=False, verbose=False,
                  deepmimic=False, backend=None):
    """sampler.py:87-130: every worker's share in one batched rollout; returns the list of all paths, worker by
    worker, as Pool.starmap would have ordered them."""
    assert mode == "samples" or mode == "trajectories"
    per = math.ceil(num_to_collect / num_workers)
    seeds = [12345 + base_seed * i for i in range(num_workers)]
    t0 = time.time()
    if getattr(policy, "eps", 0.0):
        results = []
        for i in range(num_workers):
            worker_env = copy.copy(env)
            worker_env.reset_counter = 0
            results.append(_sequential(worker_env, copy.deepcopy(policy), per, seeds[i], mode, eval_mode, deepmimic)[0])
    else:
        saved = env.reset_counter
        results, _ = _collect(env, policy, [per] * num_workers, seeds, mode, eval_mode, backend)
        env.reset_counter = saved  # the caller's env is never stepped by the workers (they get copies)
    all_paths = [p for r in results for p in r]
    if verbose:
        total = sum(len(p["rewards"]) for p in all_paths)
        print(f"Collected {total} and {len(all_paths)} trajectories in {time.time() - t0} seconds")
    return all_paths
