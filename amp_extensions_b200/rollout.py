"""Batched on-device rollout loop: policy forward + learned-dynamics step + MILO cost over a whole horizon.

Replaces the per-step host round trip of the reference's samplers (milo/milo/sampler.py:8-130 get_samples /
sample_points: `policy.get_action(o)` in numpy, `env.step(a)` at batch 1, one trajectory at a time in a
multiprocessing Pool) for E environments at once.  Every step is three library calls on device buffers —
`simstep_policy_act`, `simstep_step_cost` (or `simstep_step`), `simstep_auto_reset` — and nothing returns to the
host until the caller asks for paths or statistics.  Environments whose trajectory ends (fall contact or horizon,
gym-simenv/gym_simenv/envs/sim_env.py:164-173) start a new one on the next step, as the reference's sampler does
by calling env.reset(); the time-major buffers therefore hold several trajectories per env column, cut at the
recorded `done` flags.

The emitted paths have the dict layout mjrl/mjrl/algos/batch_reinforce.py:94-169 consumes (observations,
next_observations, actions, rewards, agent_infos{mean, log_std, evaluation}, env_infos, terminated).  With a cost
attached to the env, `rewards` already holds -cost (batch_reinforce.py:144) and the per-trajectory sums of
batch_reinforce.py:135-141 are available from `RolloutBatch.statistics()`.

Action noise: the reference draws `np.random.randn(m)` on the host per step (gaussian_mlp.py:101) after seeding
numpy per trajectory (sampler.py:38); a batched loop cannot reproduce that stream, so standard-normal draws come
from a seeded torch CUDA generator (or from the caller, `noise=`), and parity tests feed the same draws to the
oracle.
"""
import numpy as np
import torch


def _policy_parts(policy):
    """(weights, biases, nonlinearity, transformations, log_std) of an mjrl-style Gaussian MLP policy object
    (mjrl/mjrl/policies/gaussian_mlp.py:6-58 over mjrl/mjrl/utils/fc_network.py:9-41)."""
    model = policy.model
    ws = [l.weight.data for l in model.fc_layers]
    bs = [l.bias.data for l in model.fc_layers]
    nl = getattr(model, "nonlinearity", torch.tanh)
    name = "relu" if nl is torch.relu else "tanh"
    log_std = torch.as_tensor(np.asarray(policy.log_std.data if hasattr(policy.log_std, "data") else policy.log_std,
                                         dtype=np.float32)).reshape(-1)
    return ws, bs, name, (model.in_shift, model.in_scale, model.out_shift, model.out_scale), log_std


class RolloutBatch:
    """Time-major device buffers of one collect() call.  T steps x E envs."""

    def __init__(self, eng, T, E, S, A, with_cost):
        dev = eng.device
        f32 = dict(device=dev, dtype=torch.float32)
        self.eng, self.T, self.E, self.S, self.A = eng, T, E, S, A
        self.with_cost = with_cost
        self.observations = torch.empty((T, E, S), **f32)
        self.next_observations = torch.empty((T, E, S), **f32)
        self.actions = torch.empty((T, E, A), **f32)
        self.means = torch.empty((T, E, A), **f32)
        self.disc = torch.empty((T, E), **f32)
        self.done = torch.empty((T, E), device=dev, dtype=torch.uint8)
        self.cost = torch.empty((T, E), **f32) if with_cost else None
        self.ipm = torch.empty((T, E), **f32) if with_cost else None
        self.bonus = torch.empty((T, E), **f32) if with_cost else None
        self.rewards = None      # set by DeviceRollout.collect: -cost, or zeros (sim_env.py:160)
        self.final_state = torch.empty((E, S), **f32)
        self.log_std = None
        self.steps_taken = T

    # -- returns / advantages (mjrl/mjrl/utils/process_samples.py) -------------------------------
    def returns(self, gamma):
        """discount_sum of the rewards per trajectory, [T, E] (process_samples.py:3-5, 37-45)."""
        ret, _ = self.eng.discount(self.rewards, gamma, seg_end=self.done, want_returns=True, want_advantages=False)
        return ret

    def advantages(self, baseline, gamma, gae_lambda):
        """GAE advantages per trajectory against baseline [T, E] (process_samples.py:22-30): terminated
        trajectories bootstrap from 0, the unfinished trailing one from its last baseline value."""
        baseline = baseline.to(self.eng.device, torch.float32).contiguous()
        term = self.done[self.T - 1].contiguous()
        _, adv = self.eng.discount(self.rewards, gamma, baseline=baseline, gae_lambda=gae_lambda, seg_end=self.done,
                                   terminated=term, want_returns=False, want_advantages=True)
        return adv

    def normalize(self, adv, valid=None):
        """compute_advantages(normalize=True) (process_samples.py:14-19, 31-36): (adv - mean) / (std + 1e-8) with
        the mean and population std taken over every recorded step — of ALL ranks when torch.distributed is
        initialised (one all-reduce of three fp64 numbers)."""
        from . import parallel
        stats = parallel.all_reduce_sum(self.eng.moments(adv, valid))
        return self.eng.whiten(adv, stats, valid, eps=1e-8)

    # -- host views ----------------------------------------------------------------------------
    def segments(self, include_partial=True):
        """[(env, t0, t1, terminated)] with t1 exclusive, ordered by env then time."""
        done = self.done.cpu().numpy().astype(bool)
        out = []
        for e in range(self.E):
            ends = np.flatnonzero(done[:, e])
            t0 = 0
            for t in ends:
                out.append((e, t0, int(t) + 1, True))
                t0 = int(t) + 1
            if include_partial and t0 < self.T:
                out.append((e, t0, self.T, False))
        return out

    def paths(self, include_partial=True):
        """List of path dicts as milo/milo/sampler.py:70-81 builds them (float64 numpy arrays)."""
        obs = self.observations.cpu().numpy().astype(np.float64)
        nxt = self.next_observations.cpu().numpy().astype(np.float64)
        act = self.actions.cpu().numpy().astype(np.float64)
        mean = self.means.cpu().numpy()
        rew = self.rewards.cpu().numpy().astype(np.float64)
        disc = self.disc.cpu().numpy()
        log_std = np.float64(self.log_std.cpu().numpy().ravel())
        extra = {}
        if self.with_cost:
            extra = {k: getattr(self, k).cpu().numpy() for k in ("cost", "ipm", "bonus")}
        paths = []
        for (e, t0, t1, terminated) in self.segments(include_partial):
            n = t1 - t0
            infos = [{"valid": True, "disc": float(disc[t, e])} for t in range(t0, t1)]
            path = dict(observations=obs[t0:t1, e], next_observations=nxt[t0:t1, e], actions=act[t0:t1, e],
                        rewards=rew[t0:t1, e],
                        agent_infos=dict(mean=mean[t0:t1, e], log_std=np.tile(log_std, (n, 1)),
                                         evaluation=mean[t0:t1, e]),
                        env_infos=infos, terminated=terminated)
            for k, v in extra.items():
                path[k] = v[t0:t1, e]
            paths.append(path)
        return paths

    def statistics(self, include_partial=True):
        """Per-trajectory sums of batch_reinforce.py:135-141: int = -sum(bonus), ext = -sum(ipm), reward,
        ep_len; plus mean cost over all recorded steps (batch_reinforce.py:169's first term)."""
        segs = self.segments(include_partial)
        out = {"ep_len": [t1 - t0 for (_, t0, t1, _) in segs]}
        if self.with_cost:
            bonus = self.bonus.double().cpu().numpy()
            ipm = self.ipm.double().cpu().numpy()
            out["int"] = [-bonus[t0:t1, e].sum() for (e, t0, t1, _) in segs]
            out["ext"] = [-ipm[t0:t1, e].sum() for (e, t0, t1, _) in segs]
            out["reward"] = [a + b for a, b in zip(out["int"], out["ext"])]
            out["mean_cost"] = float(self.cost.double().mean().item())
        return out


class HostPathStream:
    """Double-buffered download of rollout batches for a learner on the host: the paths of collect() k cross PCIe on a
    copy stream into pinned memory while collect() k+1 already runs on the compute stream (the reference's sampler hands
    its paths over in one piece per batch as well, milo/milo/sampler.py:87-130).

        dl = HostPathStream(device)
        dl.submit(rollout.collect(T))            # returns at once
        dl.submit(rollout.collect(T))            # second horizon computes while the first downloads
        host = dl.collect()                      # pinned tensors of the OLDEST submitted batch (waits for its copy)
    """

    NAMES = ("observations", "actions", "rewards", "done", "disc")

    def __init__(self, device, names=None, depth=2):
        self.device = torch.device(device)
        self.names = tuple(names or self.NAMES)
        self.depth = int(depth)
        self.stream = torch.cuda.Stream(self.device)
        self._slots = [None] * self.depth      # pinned host tensors per slot, allocated on first use
        self._queue = []                       # (slot, event, batch) in submission order
        self._next = 0
        self.bytes_per_batch = 0

    def submit(self, batch):
        if len(self._queue) >= self.depth:
            raise RuntimeError("HostPathStream: collect() the oldest batch before submitting another")
        slot = self._next
        self._next = (self._next + 1) % self.depth
        src = {n: getattr(batch, n) for n in self.names}
        if self._slots[slot] is None or any(self._slots[slot][n].shape != t.shape for n, t in src.items()):
            self._slots[slot] = {n: torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for n, t in src.items()}
        self.bytes_per_batch = sum(t.numel() * t.element_size() for t in src.values())
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            for n, t in src.items():
                self._slots[slot][n].copy_(t, non_blocking=True)
                t.record_stream(self.stream)   # the caching allocator must not hand the buffer out before the copy
        done = torch.cuda.Event()
        done.record(self.stream)
        self._queue.append((slot, done, batch))

    def collect(self):
        if not self._queue:
            raise RuntimeError("HostPathStream: nothing submitted")
        slot, done, _ = self._queue.pop(0)
        done.synchronize()
        return self._slots[slot]


class DeviceRollout:
    """E-environment rollout loop on one GPU.

        ro = DeviceRollout(vec_env, policy, seed=0)
        batch = ro.collect(T)              # T steps of every env, auto-reset on done
        paths = batch.paths()              # what sampler.sample_points would have returned
    """

    def __init__(self, env, policy, seed=0):
        self.env = env
        self.eng = env.dynamic_ensemble.engine()
        self.device = self.eng.device
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))
        self.policy = policy
        self.load_policy(policy)
        pool = env.reset_states
        if pool is None:
            raise RuntimeError("DeviceRollout needs VecSimEnv(reset_states=...): the pool new trajectories start from")
        self.pool = torch.as_tensor(pool).to(self.device, torch.float32).contiguous()
        self._graphs = {}

    def load_policy(self, policy=None):
        """(Re-)upload the policy parameters — call after the learner's set_param_values."""
        policy = policy if policy is not None else self.policy
        ws, bs, name, (ish, isc, osh, osc), log_std = _policy_parts(policy)
        self.eng.load_policy(ws, bs, name, ish, isc, osh, osc, log_std)
        self.log_std = log_std

    def _loop(self, batch, noise, pick, T):
        env, eng = self.env, self.eng
        use_cost = batch.with_cost
        c = env.cost
        clamp = use_cost and c.cost_range is not None
        for t in range(T):
            ob = batch.observations[t]
            eng.policy_act(ob, None if noise is None else noise[t], action=batch.actions[t], mean=batch.means[t])
            if use_cost:
                eng.step_cost(ob, batch.actions[t], env.member, env.num_steps, env._w_dev, c.lambda_b,
                              env.dynamic_ensemble.threshold if clamp else 1.0, c.c_min if clamp else 0.0,
                              c.c_max if clamp else 0.0, clamp, next_state=batch.next_observations[t],
                              disc=batch.disc[t], done=batch.done[t], cost=batch.cost[t], ipm=batch.ipm[t],
                              bonus=batch.bonus[t])
            else:
                eng.step(ob, batch.actions[t], env.member, env.num_steps, next_state=batch.next_observations[t],
                         disc=batch.disc[t], done=batch.done[t])
            nxt = batch.observations[t + 1] if t + 1 < T else batch.final_state
            eng.auto_reset(batch.next_observations[t], batch.done[t], self.pool, pick[t], nxt, env.member,
                           env.num_steps)

    def _collect_graph(self, T, eval_mode, noise, pick, use_cost):
        """collect() with the whole T-step loop replayed from one CUDA graph (3 + 8 launches per step otherwise
        go through Python and the driver one by one, which dominates at small batch).  The graph and its static
        buffers are cached per configuration; the returned batch is overwritten by the next graph collect."""
        env = self.env
        E, S, A = env.num_envs, env.state_size, env.action_size
        c = env.cost
        clamp = use_cost and c.cost_range is not None
        # everything a captured launch bakes in: shapes, scalar arguments, the buffers it points at, and the
        # engine's parameter generation (load_policy / load_rff / set_rff_split re-allocate or re-plan)
        key = (T, bool(eval_mode), use_cost, self.eng.generation,
               float(env.dynamic_ensemble.threshold or 0.0) if use_cost else 0.0,
               float(c.lambda_b) if use_cost else 0.0, bool(clamp),
               (float(c.c_min), float(c.c_max)) if clamp else (0.0, 0.0),
               env._w_dev.data_ptr() if use_cost else 0, env.ob.data_ptr(), env.member.data_ptr(),
               env.num_steps.data_ptr(), self.pool.data_ptr())
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs.clear()  # one entry: a stale graph pins a whole T x E batch and may point at freed memory
            batch = RolloutBatch(self.eng, T, E, S, A, use_cost)
            s_noise = None if eval_mode else torch.zeros((T, E, A), device=self.device, dtype=torch.float32)
            s_pick = torch.zeros((T, E), device=self.device, dtype=torch.int32)
            # warm-up outside capture (workspace allocation, kernel attributes), on copies of the env's counters
            saved = (env.ob.clone(), env.member.clone(), env.num_steps.clone())
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                batch.observations[0].copy_(env.ob)
                self._loop(batch, s_noise, s_pick, min(T, 2))
            torch.cuda.current_stream(self.device).wait_stream(side)
            env.ob.copy_(saved[0]); env.member.copy_(saved[1]); env.num_steps.copy_(saved[2])
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                batch.observations[0].copy_(env.ob)
                self._loop(batch, s_noise, s_pick, T)
                env.ob.copy_(batch.final_state)
            ent = (g, batch, s_noise, s_pick)
            self._graphs[key] = ent
        g, batch, s_noise, s_pick = ent
        if s_noise is not None:
            if noise is not None:
                s_noise.copy_(noise)
            else:
                s_noise.normal_(generator=self.gen)
        if pick is not None:
            s_pick.copy_(pick)
        else:
            s_pick.random_(0, self.pool.shape[0], generator=self.gen)
        g.replay()
        batch.rewards = -batch.cost if use_cost else torch.zeros((T, E), device=self.device, dtype=torch.float32)
        batch.log_std = self.log_std
        return batch

    def collect(self, num_steps, eval_mode=False, noise=None, pick=None, batch=None, graph=False):
        """Run `num_steps` steps of every env from the env's current state.  eval_mode: the mean action is used
        (sampler.py:51).  noise [T, E, A] / pick [T, E] int32 override the generator (tests).  graph=True replays
        the loop from a cached CUDA graph (the returned batch is then reused by the next such call)."""
        env = self.env
        T, E, S, A = int(num_steps), env.num_envs, env.state_size, env.action_size
        use_cost = env.cost is not None and env._w_dev is not None
        if graph:
            return self._collect_graph(T, eval_mode, noise, pick, use_cost)
        if batch is None:
            batch = RolloutBatch(self.eng, T, E, S, A, use_cost)
        if noise is None and not eval_mode:
            noise = torch.randn((T, E, A), device=self.device, dtype=torch.float32, generator=self.gen)
        if eval_mode:
            noise = None
        if pick is None:
            pick = torch.randint(0, self.pool.shape[0], (T, E), device=self.device, dtype=torch.int32,
                                 generator=self.gen)
        batch.observations[0].copy_(env.ob)
        self._loop(batch, noise, pick, T)
        env.ob.copy_(batch.final_state)
        batch.rewards = -batch.cost if use_cost else torch.zeros((T, E), device=self.device, dtype=torch.float32)
        batch.log_std = self.log_std
        return batch

    def sample_paths(self, num_to_collect, mode="samples", eval_mode=False, max_steps=None):
        """sampler.sample_points's contract: at least `num_to_collect` samples (mode 'samples') or trajectories
        (mode 'trajectories'), complete trajectories only (sampler.py:30-34, 79)."""
        assert mode in ("samples", "trajectories")
        env = self.env
        env.reset()
        paths, n_samples = [], 0
        horizon = env.horizon
        while True:
            need = num_to_collect - (n_samples if mode == "samples" else len(paths))
            if need <= 0:
                break
            T = max_steps or horizon
            batch = self.collect(T, eval_mode=eval_mode)
            new = batch.paths(include_partial=False)
            paths.extend(new)
            n_samples += sum(len(p["rewards"]) for p in new)
            # unfinished trailing segments are dropped, like a worker that stops after its last full trajectory
            env.reset()
        return paths, n_samples
