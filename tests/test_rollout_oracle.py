"""oracle/rollout_oracle.py against the reference's own mjrl outputs (tests/golden/rollout_golden.npz, written by
tests/golden/make_rollout_golden.py from /root/reference) — CPU only."""
import os

import numpy as np
import torch

from oracle import rollout_oracle as ro

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_golden.npz"))


def _policy(tag, n_layers):
    ws = [torch.from_numpy(G[f"{tag}/w{i}"]) for i in range(n_layers)]
    bs = [torch.from_numpy(G[f"{tag}/b{i}"]) for i in range(n_layers)]
    return ws, bs


def test_fc_forward_default_transformations_is_bit_exact():
    ws, bs = _policy("polA", 3)
    mean = ro.fc_forward(ws, bs, G["polA/obs"]).numpy()
    assert np.array_equal(mean, G["polA/mean"])


def test_fc_forward_relu_with_transformations_is_bit_exact():
    ws, bs = _policy("polB", 4)
    mean = ro.fc_forward(ws, bs, G["polB/obs"], "relu", G["polB/in_shift"], G["polB/in_scale"], G["polB/out_shift"],
                         G["polB/out_scale"]).numpy()
    assert np.array_equal(mean, G["polB/mean"])


def test_get_action_reproduces_the_reference_stream():
    a = ro.get_action(G["polA/get_action_mean"], G["polA/log_std"], G["polA/noise"])
    assert a.dtype == np.float64
    assert np.array_equal(a, G["polA/get_action"])
    # the batch-1 forward get_action runs agrees with the batched forward to fp32 round-off
    assert np.abs(G["polA/get_action_mean"] - G["polA/mean"][:8]).max() < 1e-6


def test_returns_and_gae_match_process_samples():
    gamma, lam = G["ps/gamma_lambda"]
    for i, (n, term) in enumerate(zip(G["ps/lens"], G["ps/terminated"])):
        r, b = G[f"ps/rewards{i}"], G[f"ps/baseline{i}"]
        assert len(r) == n
        assert np.array_equal(ro.discount_sum(r, gamma), G[f"ps/returns{i}"])
        assert np.array_equal(ro.gae_advantages(r, b, bool(term), gamma, lam), G[f"ps/advantages{i}"])


# ---- the whole loop against the reference's own sampler.get_samples (tests/golden/sampler_golden.npz) ----------

def sampler_golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "sampler_golden.npz"))


def sampler_noise(g):
    """The draws MLP.get_action made inside get_samples: np.random.seed(seed + k) per trajectory (sampler.py:36-38),
    then uniform() and randn(action_dim) per step (gaussian_mlp.py:95-104)."""
    n_traj, T, A = g["actions"].shape
    noise = np.zeros((T, n_traj, A))
    for k in range(n_traj):
        rs = np.random.RandomState(int(g["seed"]) + k + 1)
        for t in range(int(g["length"][k])):
            rs.uniform()
            noise[t, k] = rs.randn(A)
    return noise


def test_rollout_loop_matches_the_reference_sampler():
    """get_samples + SimEnv + DynamicsEnsemble + MLP policy of the reference, six trajectories (early falls and horizon
    cuts), replayed by the batched restatement with the same exploration draws: observations, actions, policy means,
    trajectory lengths and the member round-robin."""
    from oracle import milo_oracle as mo
    from tests import helpers as H
    g = sampler_golden()
    N, hidden, horizon = int(g["N"]), [int(h) for h in g["hidden"]], int(g["horizon"])
    s, a, s2 = H.synth_dataset(int(g["dataset_rows"]), 226, 28, int(g["dataset_seed"]))
    ws, bs = mo.init_ensemble(226, 28, hidden, N, base_seed=int(g["base_seed"]), dense_connect=True)
    np.testing.assert_allclose([[float(w.double().abs().sum()) for w in m] for m in ws], g["weight_checksum"], rtol=1e-12)
    tf = mo.get_transformations(s, a, s2)
    policy = dict(ws=[torch.from_numpy(g[f"pol_w{i}"]) for i in range(3)], bs=[torch.from_numpy(g[f"pol_b{i}"]) for i in range(3)],
                  log_std=g["log_std"])
    n_traj, T, _ = g["actions"].shape
    noise = sampler_noise(g)
    res = ro.rollout(ws, bs, tf, policy, g["observations"][:, 0], g["member"], np.zeros(n_traj, dtype=np.int64),
                     g["observations"][:, 0], noise, np.zeros((T, n_traj), dtype=np.int64), horizon=horizon, n_models=N)
    for k in range(n_traj):
        n = int(g["length"][k])
        first_done = int(np.argmax(res["done"][:, k])) + 1
        assert res["done"][:, k].any() and first_done == n, (k, first_done, n)     # same trajectory length
        np.testing.assert_allclose(res["means"][:n, k], g["means"][k, :n], atol=2e-5)
        np.testing.assert_allclose(res["actions"][:n, k], g["actions"][k, :n], atol=2e-5)
        np.testing.assert_allclose(res["observations"][:n, k], g["observations"][k, :n], atol=2e-5)
        np.testing.assert_allclose(res["next_observations"][:n, k], g["next_observations"][k, :n], atol=2e-5)
    assert sorted(set(int(x) for x in g["length"])) != [horizon]   # some trajectories ended on a fall


# ---- BatchREINFORCE.train_step's reward replacement against the reference's own run (trainstep_golden.npz) -----

class _OracleEnsemble:
    def __init__(self, ws, bs, tf, threshold):
        self.ws, self.bs, self.tf, self.threshold = ws, bs, tf, threshold

    def get_action_discrepancy(self, s, a):   # DYN:154-165
        from oracle import milo_oracle as mo
        return mo.discrepancy_from_preds(mo.ensemble_forward(self.ws, self.bs, self.tf, s, a))


def test_reward_replacement_matches_the_reference_train_step():
    """tests/golden/make_trainstep_golden.py ran the reference's BatchREINFORCE.train_step (fit_cost on the rollout,
    per-trajectory get_bonus_costs, reward = -cost, int / ext / ep_len sums, mb_mmd, bonus_mmd, then compute_returns
    and compute_advantages) over the reference RBFLinearCost and DynamicsEnsemble.  The same lines
    (tests.helpers.reward_replacement) over the oracle objects must reproduce every number."""
    from oracle import milo_oracle as mo
    from tests import helpers as H
    g = H.trainstep_golden()
    N, hidden = int(g["N"]), [int(h) for h in g["hidden"]]
    s, a, s2 = H.synth_dataset(int(g["dataset_rows"]), 226, 28, int(g["dataset_seed"]))
    ws, bs = mo.init_ensemble(226, 28, hidden, N, base_seed=int(g["base_seed"]), dense_connect=True)
    tf = mo.get_transformations(s, a, s2)
    threshold = mo.compute_threshold(ws, bs, tf, s, a, batch_size=256)   # over the offline dataset, DYN:145-152
    assert abs(threshold - float(g["threshold"])) < 1e-6 * float(g["threshold"])

    class Cost(mo.RffCostOracle):
        def get_bonus_costs(self, states, actions, ensemble, next_states=None):   # the reference signature (LC:111)
            return mo.RffCostOracle.get_bonus_costs(self, states, actions, ensemble.get_action_discrepancy(states, actions),
                                                    ensemble.threshold, next_states=next_states)

    cost = Cost(torch.from_numpy(g["expert"]), feature_dim=int(g["feature_dim"]), input_type="ss",
                bw_quantile=float(g["bw_quantile"]), lambda_b=float(g["lambda_b"]), seed=int(g["cost_seed"]))
    assert cost.bw == float(g["cost_bw"])
    paths = H.trainstep_paths(g)
    infos = H.reward_replacement(paths, cost, _OracleEnsemble(ws, bs, tf, float(g["threshold"])))
    np.testing.assert_allclose(cost.w.numpy(), g["cost_w"], atol=1e-6)
    assert abs(infos["mb_mmd"] - float(g["mb_mmd"])) < 1e-6
    assert abs(infos["bonus_mmd"] - float(g["bonus_mmd"])) < 1e-6
    assert infos["ep_len"] == list(g["info_ep_len"])
    for key in ("int", "ext", "reward"):
        np.testing.assert_allclose(infos[key], g[f"info_{key}"], atol=1e-5)
    for k, p in enumerate(paths):
        n = len(p["rewards"])
        np.testing.assert_allclose(p["rewards"], g["rewards"][k, :n], atol=1e-6)
        np.testing.assert_allclose(ro.discount_sum(p["rewards"], float(g["gamma"])), g["returns"][k, :n], atol=1e-6)
        adv = ro.gae_advantages(p["rewards"], np.zeros(n), True, float(g["gamma"]), float(g["gae_lambda"]))
        np.testing.assert_allclose(adv, g["advantages"][k, :n], atol=1e-6)
    assert g["rewards"].min() < 0 < g["rewards"].max()   # both the pessimism bonus and the IPM term are active
