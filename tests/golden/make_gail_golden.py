"""Generates tests/golden/gail_golden.npz by running the REFERENCE's own GAILCost (milo/milo/gail_cost.py).
Build container only (needs /root/reference):

    python tests/golden/make_gail_golden.py

gail_cost.py imports `milo.datasets`; the milo package's __init__ needs gym, so a bare `milo` package module is
registered whose `datasets` submodule is the reference file loaded by path.  Pins oracle/milo_oracle.py::gail_* and
amp_extensions_b200.GAILCost.  Nothing at test time reads /root/reference.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SIMSTEP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gail_golden.npz")


def load_gail():
    pkg = types.ModuleType("milo")
    pkg.__path__ = []
    sys.modules["milo"] = pkg
    for name in ("datasets", "gail_cost"):
        spec = importlib.util.spec_from_file_location(f"milo.{name}", os.path.join(REF, "milo", "milo", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"milo.{name}"] = m
        spec.loader.exec_module(m)
        setattr(pkg, name, m)
    return sys.modules["milo.gail_cost"]


class _Ens:
    def __init__(self, disc):
        self.disc = disc

    def get_action_discrepancy(self, states, actions):
        return self.disc.clone()


def main():
    gc = load_gail()
    out = {"torch_version": np.array(torch.__version__)}
    S, A = 226, 28
    g = torch.Generator().manual_seed(4)
    es = torch.randn(160, S, generator=g)
    expert = torch.cat([es, es + 0.05 * torch.randn(160, S, generator=g)], dim=1)
    xs, xa = torch.randn(96, S, generator=g), torch.randn(96, A, generator=g)
    nxt = xs + 0.05 * torch.randn(96, S, generator=g)
    disc = torch.rand(96, generator=g) * 0.6
    out["expert"], out["xs"], out["xa"], out["next"], out["disc"] = (expert.numpy(), xs.numpy(), xa.numpy(), nxt.numpy(),
                                                                     disc.numpy())
    cases = {"ls_two_hidden": dict(hidden_dims=[192, 128], disc_loss_type="least_squares", lambda_b=0.5),
             "ll_small": dict(hidden_dims=[96, 40], disc_loss_type="log_likelihood", lambda_b=0.2),
             "ls_linear": dict(hidden_dims=[], disc_loss_type="least_squares", lambda_b=0.7)}
    for tag, kw in cases.items():
        cost = gc.GAILCost(expert, None, feature_dim=1, input_type="ss", seed=100, **kw)
        # a few discriminator updates so the outputs are those of a trained net, not of the initialisation
        for it in range(3):
            np.random.seed(10 + it)
            cost.update_disc(torch.cat([xs, nxt], dim=1)[: 64].clone())
        net = cost.disc.net
        lin = [net] if isinstance(net, torch.nn.Linear) else [m for m in net if isinstance(m, torch.nn.Linear)]
        out[f"{tag}/n_linear"] = np.array(len(lin))
        out[f"{tag}/loss_type"] = np.array(kw["disc_loss_type"])
        out[f"{tag}/lambda_b"] = np.array(kw["lambda_b"])
        for i, l in enumerate(lin):
            out[f"{tag}/w{i}"], out[f"{tag}/b{i}"] = l.weight.data.numpy().copy(), l.bias.data.numpy().copy()
        ss = torch.cat([xs, nxt], dim=1)
        with torch.no_grad():
            out[f"{tag}/disc_outs"] = cost.disc(ss).numpy()
        out[f"{tag}/costs"] = cost.get_costs(ss).numpy()
        total, info = cost.get_bonus_costs(xs, xa, _Ens(disc), next_states=nxt)
        out[f"{tag}/total"] = total.numpy()
        for k in ("bonus", "ipm", "v_targ", "cost"):
            out[f"{tag}/info_{k}"] = info[k].numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    main()
