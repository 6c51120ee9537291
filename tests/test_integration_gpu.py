"""The pieces together: one MILO iteration (examples/milo_iteration.py) with the reference's objects swapped for this
package's — ensemble training, threshold, IPM cost, on-device rollout from clip resets, cost fit, returns / GAE /
whitening / statistics — checked for the invariants the reference's trainer relies on."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_one_milo_iteration_end_to_end():
    import milo_iteration
    out, batch = milo_iteration.main(num_envs=128, horizon=24, epochs=2, n_offline=1024, hidden=(64, 64), num_models=3,
                                     verbose=False)
    assert all(b <= f for b, f in zip(out["train_loss_best"], out["train_loss_first"]))   # training went downhill
    assert out["threshold"] > 0 and out["mmd"] >= 0
    assert out["env_steps"] == 128 * 24 and out["trajectories"] >= 128
    assert -1.0 - 1e-6 <= out["mean_cost"] <= 0.0025 + 1e-6          # (1 - lambda) c - lambda * bonus, c in [-1, 0]
    assert abs(out["adv_mean"]) < 1e-4 and abs(out["adv_std"] - 1.0) < 1e-3               # whitened
    assert np.isfinite(out["mean_return"]) and np.isfinite(out["expert_cost"])
    # rewards are -cost (batch_reinforce.py:144) and every path is a contiguous slice of one env's column
    assert torch.equal(batch.rewards, -batch.cost)
    paths = batch.paths()
    assert sum(len(p["rewards"]) for p in paths) == out["env_steps"]
    assert all(p["observations"].shape[1] == 226 and p["actions"].shape[1] == 28 for p in paths)
