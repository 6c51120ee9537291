"""Generates tests/golden/mlpcost_golden.npz by running the REFERENCE's own MLPCost
(milo/milo/linear_cost.py:154-301).  Build container only (needs /root/reference):

    python tests/golden/make_mlpcost_golden.py

Pins oracle/milo_oracle.py::MlpCostOracle (tests/test_oracle.py) and amp_extensions_b200.MLPCost
(tests/test_parity_gpu.py).  Nothing at test time reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mlpcost_golden.npz")


class _Ens:
    def __init__(self, disc, threshold):
        self.disc, self.threshold = disc, threshold

    def get_action_discrepancy(self, states, actions):
        return self.disc.clone()


def main():
    ref = load_reference()
    out = {"torch_version": np.array(torch.__version__)}
    S, A = 226, 28
    g = torch.Generator().manual_seed(2)
    es = torch.randn(192, S, generator=g)
    expert = torch.cat([es, es + 0.05 * torch.randn(192, S, generator=g)], dim=1)
    xs, xa = torch.randn(80, S, generator=g), torch.randn(80, A, generator=g)
    nxt = xs + 0.05 * torch.randn(80, S, generator=g)
    disc = torch.rand(80, generator=g) * 0.6
    out["expert"], out["xs"], out["xa"], out["next"], out["disc"] = (expert.numpy(), xs.numpy(), xa.numpy(), nxt.numpy(),
                                                                     disc.numpy())
    cases = {"two_hidden": dict(hidden_dims=[96, 160], activation="relu", feature_dim=72),
             "one_hidden_quirk": dict(hidden_dims=[64], activation="relu", feature_dim=64),
             "three_tanh": dict(hidden_dims=[48, 40, 56], activation="tanh", feature_dim=40)}
    for tag, kw in cases.items():
        cost = ref["linear_cost"].MLPCost(expert, input_type="ss", bw_quantile=0.1, lambda_b=0.3, seed=100, **kw)
        lin = [m for m in cost.net if isinstance(m, torch.nn.Linear)]
        out[f"{tag}/n_linear"] = np.array(len(lin))
        out[f"{tag}/hidden"] = np.array(kw["hidden_dims"])
        out[f"{tag}/act"] = np.array(kw["activation"])
        out[f"{tag}/feature_dim"] = np.array(kw["feature_dim"])
        for i, l in enumerate(lin):
            out[f"{tag}/w{i}"], out[f"{tag}/b{i}"] = l.weight.data.numpy(), l.bias.data.numpy()
        pi = torch.cat([xs, nxt], dim=1)
        out[f"{tag}/bw"] = np.array(cost.bw)
        out[f"{tag}/phi_e"] = cost.phi_e.numpy()
        out[f"{tag}/rep"] = cost.get_rep(pi).numpy()
        out[f"{tag}/mmd"] = np.array(cost.fit_cost(pi))
        out[f"{tag}/w"] = cost.w.numpy()
        out[f"{tag}/costs"] = cost.get_costs(pi).numpy()
        out[f"{tag}/expert_cost"] = np.array(float(cost.get_expert_cost()))
        total, info = cost.get_bonus_costs(xs, xa, _Ens(disc, 0.4), next_states=nxt)
        out[f"{tag}/total"] = total.numpy()
        for k in ("bonus", "ipm", "v_targ", "cost"):
            out[f"{tag}/info_{k}"] = info[k].numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    main()
