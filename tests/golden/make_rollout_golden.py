"""Generates tests/golden/rollout_golden.npz by running the REFERENCE's own mjrl code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_rollout_golden.py

Imports the reference's mjrl/mjrl/utils/fc_network.py, mjrl/mjrl/utils/process_samples.py and
mjrl/mjrl/policies/gaussian_mlp.py (the mjrl package root is put on sys.path; its __init__ files import
nothing), evaluates them on seeded inputs and stores inputs + outputs.  The fixtures pin
oracle/rollout_oracle.py (tests/test_rollout_oracle.py) and the CUDA policy / discount kernels
(tests/test_rollout_gpu.py).  Nothing at test time reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SIMSTEP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rollout_golden.npz")


class _Baseline:
    def __init__(self, values):
        self.values = values

    def predict(self, path):
        return self.values[path["id"]]


def main():
    sys.path.insert(0, os.path.join(REF, "mjrl"))
    from mjrl.policies.gaussian_mlp import MLP
    from mjrl.utils import process_samples as ps
    from mjrl.utils.fc_network import FCNetwork

    out = {"torch_version": np.array(torch.__version__)}

    # ---- policy A: MILO's humanoid policy shape (226 -> 32 -> 32 -> 28, tanh), default transformations ----
    pol = MLP(226, 28, hidden_sizes=(32, 32), seed=123, init_log_std=-0.5, min_log_std=-2.5)
    rng = np.random.RandomState(7)
    obs = rng.randn(96, 226).astype(np.float32) * 1.5
    with torch.no_grad():
        mean = pol.model(torch.from_numpy(obs)).numpy()
    for i, l in enumerate(pol.model.fc_layers):
        out[f"polA/w{i}"], out[f"polA/b{i}"] = l.weight.data.numpy(), l.bias.data.numpy()
    out["polA/log_std"] = pol.log_std.data.numpy()
    out["polA/obs"], out["polA/mean"] = obs, mean
    # get_action with a known numpy stream (gaussian_mlp.py:95-104): eps = 0 draws uniform() first, then randn(m)
    np.random.seed(11)
    acts, noises, means1 = [], [], []
    for i in range(8):
        st = np.random.get_state()
        a, info = pol.get_action(obs[i])
        np.random.set_state(st)
        np.random.uniform()
        noises.append(np.random.randn(28))
        acts.append(a)
        means1.append(info["mean"])  # batch-1 forward: may differ from the batched sgemm in the last ulp
        assert np.allclose(info["evaluation"], mean[i], atol=1e-6)
    out["polA/get_action"], out["polA/noise"] = np.array(acts), np.array(noises)
    out["polA/get_action_mean"] = np.array(means1)

    # ---- policy B: relu, three hidden layers, all four transformations set ----
    torch.manual_seed(5)
    in_shift, in_scale = rng.randn(40).astype(np.float32), (0.5 + rng.rand(40)).astype(np.float32)
    out_shift, out_scale = rng.randn(12).astype(np.float32), (0.5 + rng.rand(12)).astype(np.float32)
    net = FCNetwork(40, 12, hidden_sizes=(64, 48, 33), nonlinearity="relu", in_shift=in_shift, in_scale=in_scale,
                    out_shift=out_shift, out_scale=out_scale)
    obs_b = rng.randn(50, 40).astype(np.float32)
    with torch.no_grad():
        mean_b = net(torch.from_numpy(obs_b)).numpy()
    for i, l in enumerate(net.fc_layers):
        out[f"polB/w{i}"], out[f"polB/b{i}"] = l.weight.data.numpy(), l.bias.data.numpy()
    out["polB/in_shift"], out["polB/in_scale"] = in_shift, in_scale
    out["polB/out_shift"], out["polB/out_scale"] = out_shift, out_scale
    out["polB/obs"], out["polB/mean"] = obs_b, mean_b

    # ---- process_samples: returns and GAE advantages over ragged paths ----
    lens = [1, 2, 17, 300, 64, 5]
    term = [True, False, True, True, False, False]
    paths, base = [], []
    for i, (n, tm) in enumerate(zip(lens, term)):
        paths.append(dict(id=i, rewards=rng.randn(n), terminated=tm))
        base.append(rng.randn(n))
    gamma, lam = 0.995, 0.97
    ps.compute_returns(paths, gamma)
    ps.compute_advantages(paths, _Baseline(base), gamma, lam)
    out["ps/lens"], out["ps/terminated"] = np.array(lens), np.array(term)
    out["ps/gamma_lambda"] = np.array([gamma, lam])
    for i, p in enumerate(paths):
        out[f"ps/rewards{i}"], out[f"ps/baseline{i}"] = p["rewards"], base[i]
        out[f"ps/returns{i}"], out[f"ps/advantages{i}"] = p["returns"], p["advantages"]
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in list(out.items())[:6]})


if __name__ == "__main__":
    main()
