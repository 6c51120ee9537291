"""Writes tests/golden/trainstep_golden.npz by running the REFERENCE's own mjrl BatchREINFORCE.train_step up to and
including its reward replacement (mjrl/mjrl/algos/batch_reinforce.py:85-180) on the CPU:

    paths  <- milo/milo/sampler.py::get_samples over the reference SimEnv + DynamicsEnsemble + MLP policy
    fit    <- RBFLinearCost.fit_cost on the rollout (batch_reinforce.py:107-115)
    reward <- -RBFLinearCost.get_bonus_costs(states, actions, ensemble, next_states) per trajectory (:119-144)
    infos  <- int / ext / reward / ep_len per trajectory, mb_mmd, bonus_mmd (:135-141, :169)
    then process_samples.compute_returns / compute_advantages (:178-180)

Run in the build container only (needs /root/reference):  python tests/golden/make_trainstep_golden.py

Stubs: gym and the SWIG simulator (as make_simenv_golden.py), tkinter (dynamics.py), `mjrl.samplers.core` (imports a
gym wrapper; train_step does not use it in 'model_based' mode), `mjrl.utils.logger` (imports matplotlib; unused with
save_logs=False).  batch_reinforce.sample_points is pointed at a
function that returns the paths collected by get_samples in this process (the reference forks a worker pool there),
and train_from_paths / the baseline, which belong to the NPG update and not to this path, are replaced by recorders.
"""
import copy
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import make_golden as mg  # noqa: E402
import make_simenv_golden as msg  # noqa: E402
from oracle import imitation_oracle as io  # noqa: E402
from tests import helpers as H  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trainstep_golden.npz")
S, A = msg.S, msg.A


class ZeroBaseline:
    def predict(self, path):
        return np.zeros(len(path["rewards"]))

    def fit(self, paths, return_errors=False, return_all_errors=False):
        return 0.0, 0.0, []


def main():
    ref = mg.load_reference()
    SimEnv = msg.load_reference_simenv()
    sys.path.insert(0, os.path.join(mg.REF, "mjrl"))
    spec = importlib.util.spec_from_file_location("ref_sampler", os.path.join(mg.REF, "milo", "milo", "sampler.py"))
    sampler = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sampler)
    import mjrl.samplers  # the real (empty) package; only its gym-dependent `core` module is stubbed
    core = types.ModuleType("mjrl.samplers.core")
    sys.modules["mjrl.samplers.core"] = core
    mjrl.samplers.core = core
    import mjrl.utils
    logger = types.ModuleType("mjrl.utils.logger")   # imports matplotlib (absent); DataLog is unused with save_logs=False
    logger.DataLog = object
    sys.modules["mjrl.utils.logger"] = logger
    mjrl.utils.logger = logger
    milo_pkg, milo_sampler = types.ModuleType("milo"), types.ModuleType("milo.sampler")
    milo_pkg.__path__ = []
    milo_sampler.sample_points = sampler.sample_points
    sys.modules.update({"milo": milo_pkg, "milo.sampler": milo_sampler})
    from mjrl.algos import batch_reinforce as br
    from mjrl.policies.gaussian_mlp import MLP
    msg.FakeSimulator.clip = io.Clip(H.spinkick_raw(), io.HUMANOID3D, "wrap")
    os.chdir(msg.DM_ROOT)

    N, hidden, horizon, seed, n_traj = 3, [32, 32], 8, 70, 10
    s, a, s2 = mg.synth_dataset(512, S, A, 0)
    ds = ref["datasets"].AmpDataset(s, a, s2)
    ens = mg.build_ensemble(ref, S, A, N, hidden, True, "relu", ds)
    ens.compute_threshold()   # maximum discrepancy over the offline dataset's own dataloader (dynamics.py:145-152)
    clip = msg.FakeSimulator.clip
    ts = np.linspace(0.0, clip.duration, 128, endpoint=False)
    th_s = torch.from_numpy(np.stack([io.record_state(io.HUMANOID3D, clip.kin_pose(t), clip.kin_vel(t)) for t in ts])).float()
    env = SimEnv(ens, deepmimic_args=msg.ARG_FILE, horizon=horizon, seed=1)
    pol = MLP(S, A, hidden_sizes=(32, 32), seed=123, init_log_std=-1.0, min_log_std=-2.5)
    expert = torch.cat([th_s[:96], th_s[1:97]], dim=1)      # consecutive clip states as the expert (s, s') pairs
    cost = ref["linear_cost"].RBFLinearCost(expert, feature_dim=64, input_type="ss", bw_quantile=0.1, lambda_b=0.1, seed=100)

    paths, _ = sampler.get_samples(env, pol, n_traj, seed, mode="trajectories", eval_mode=False)
    raw = copy.deepcopy(paths)
    for p in paths:   # what a real DeepMimic env adds and train_step(deepmimic=True) expects; unused afterwards
        p["env_infos"] = [{"valid": True} for _ in p["env_infos"]]
    br.sample_points = lambda **kw: paths
    agent = br.BatchREINFORCE(env, pol, ZeroBaseline(), seed=seed, save_logs=False)
    seen = {}

    def record(paths_, infos_):
        seen["paths"], seen["infos"] = paths_, infos_
        return []

    agent.train_from_paths = record
    reward_kwargs = dict(reward_func=cost, ensemble=ens, gail_cost=False, device=torch.device("cpu"))
    agent.train_step(N=n_traj, env=env, sample_mode="model_based", cost_input_type="ss", gamma=0.99, gae_lambda=0.95,
                     num_cpu=1, num_samples=1, reward_kwargs=reward_kwargs)
    out_paths, infos = seen["paths"], seen["infos"]

    T = max(len(p["rewards"]) for p in raw)
    obs, nxt, act = np.zeros((n_traj, T, S)), np.zeros((n_traj, T, S)), np.zeros((n_traj, T, A))
    rew, ret, adv = np.zeros((n_traj, T)), np.zeros((n_traj, T)), np.zeros((n_traj, T))
    length = np.zeros(n_traj, dtype=np.int64)
    for k, (p0, p1) in enumerate(zip(raw, out_paths)):
        n = len(p0["rewards"])
        length[k] = n
        obs[k, :n], nxt[k, :n], act[k, :n] = p0["observations"], p0["next_observations"], p0["actions"]
        rew[k, :n], ret[k, :n], adv[k, :n] = p1["rewards"], p1["returns"], p1["advantages"]
    out = dict(N=np.int64(N), hidden=np.array(hidden), horizon=np.int64(horizon), dataset_seed=np.int64(0),
               dataset_rows=np.int64(512), base_seed=np.int64(100), threshold=np.float64(ens.threshold),
               expert=expert.numpy(),
               feature_dim=np.int64(64), bw_quantile=np.float64(0.1), lambda_b=np.float64(0.1), cost_seed=np.int64(100),
               gamma=np.float64(0.99), gae_lambda=np.float64(0.95),
               observations=obs, next_observations=nxt, actions=act, length=length,
               rewards=rew, returns=ret, advantages=adv,
               info_int=np.array(infos["int"]), info_ext=np.array(infos["ext"]), info_reward=np.array(infos["reward"]),
               info_ep_len=np.array(infos["ep_len"]), mb_mmd=np.float64(infos["mb_mmd"]),
               bonus_mmd=np.float64(infos["bonus_mmd"]), cost_w=cost.w.numpy().copy(), cost_bw=np.float64(cost.bw),
               weight_checksum=np.array([[float(l.weight.detach().double().abs().sum()) for l in m.model.fc_layers]
                                         for m in ens.models]))
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes; lengths", length, "threshold", ens.threshold, "mb_mmd", infos["mb_mmd"],
          "bonus_mmd", float(infos["bonus_mmd"]), "reward range", rew.min(), rew.max())


if __name__ == "__main__":
    main()
