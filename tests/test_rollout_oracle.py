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
