"""Generates tests/golden/milo_golden.npz by running the REFERENCE's own code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports the reference's milo/milo/{datasets,dynamics,linear_cost}.py by file path (a two-line tkinter stub
is needed because dynamics.py:2 imports tkinter.messagebox), evaluates them on seeded synthetic inputs and
stores inputs + outputs.  The fixtures pin oracle/milo_oracle.py (tests/test_oracle.py) and, through it and
directly, the CUDA path (tests/test_parity_gpu.py).  Nothing at test time reads /root/reference.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SIMSTEP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "milo_golden.npz")


def load_reference():
    tk = types.ModuleType("tkinter")
    mb = types.ModuleType("tkinter.messagebox")
    mb.NO = "no"
    tk.messagebox = mb
    tk.E = "e"
    sys.modules.setdefault("tkinter", tk)
    sys.modules.setdefault("tkinter.messagebox", mb)
    mods = {}
    for name in ("datasets", "dynamics", "linear_cost"):
        spec = importlib.util.spec_from_file_location(f"ref_milo_{name}", os.path.join(REF, "milo", "milo", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def synth_dataset(M, S, A, seed):
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(M, S, generator=g)
    a = torch.randn(M, A, generator=g)
    s2 = s + 0.05 * torch.randn(M, S, generator=g)
    return s, a, s2


def build_ensemble(ref, S, A, N, hidden, dense, act, ds, base_seed=100):
    ens = ref["dynamics"].DynamicsEnsemble(S, A, ds, None, num_models=N, batch_size=256, hidden_sizes=hidden,
                                           dense_connect=dense, activation=act, transform=True, base_seed=base_seed,
                                           num_workers=0)
    for m in ens.models:  # what load_ensemble does (dynamics.py:128-131)
        (m.state_mean, m.state_scale, m.action_mean, m.action_scale, m.diff_mean, m.diff_scale) = ens.transformations
    return ens


def main():
    ref = load_reference()
    out = {"torch_version": np.array(torch.__version__)}

    # ---- case "tiny": explicit weights stored, several architectures --------------------------------
    for tag, (S, A, N, hidden, dense, act) in {
        "tiny_dense": (20, 6, 3, [32, 24], True, "relu"),
        "tiny_plain_tanh": (20, 6, 2, [16, 16], False, "tanh"),
    }.items():
        s, a, s2 = synth_dataset(512, S, A, seed=0)
        ds = ref["datasets"].AmpDataset(s, a, s2)
        ens = build_ensemble(ref, S, A, N, hidden, dense, act, ds)
        g = torch.Generator().manual_seed(1)
        xs, xa = torch.randn(64, S, generator=g), torch.randn(64, A, generator=g)
        with torch.no_grad():
            preds = torch.stack([m.forward(xs, xa) for m in ens.models])
            preds_norm = torch.stack([m.forward(xs, xa, unnormalize_out=False) for m in ens.models])
        disc = ens.compute_discrepancy(xs, xa)
        torch.manual_seed(5)
        ens.compute_threshold()
        out[f"{tag}/dims"] = np.array([S, A, N, int(dense)] + hidden)
        out[f"{tag}/act"] = np.array(act)
        out[f"{tag}/ds_s"], out[f"{tag}/ds_a"], out[f"{tag}/ds_s2"] = s.numpy(), a.numpy(), s2.numpy()
        for i, t in enumerate(ens.transformations):
            out[f"{tag}/tf{i}"] = t.numpy()
        for k, m in enumerate(ens.models):
            for name, v in m.model.state_dict().items():
                out[f"{tag}/m{k}/{name}"] = v.numpy()
        out[f"{tag}/xs"], out[f"{tag}/xa"] = xs.numpy(), xa.numpy()
        out[f"{tag}/preds"], out[f"{tag}/preds_norm"] = preds.numpy(), preds_norm.numpy()
        out[f"{tag}/disc"] = disc.numpy()
        out[f"{tag}/threshold"] = np.array(ens.threshold)

    # ---- case "ns": the north-star ensemble, 4 x (512 x 4) dense-connect, humanoid3d dims -------------
    S, A, N, hidden = 226, 28, 4, [512] * 4
    s, a, s2 = synth_dataset(8192, S, A, seed=0)
    ds = ref["datasets"].AmpDataset(s, a, s2)
    ens = build_ensemble(ref, S, A, N, hidden, True, "relu", ds)
    g = torch.Generator().manual_seed(1)
    xs, xa = torch.randn(48, S, generator=g), torch.randn(48, A, generator=g)
    with torch.no_grad():
        preds = torch.stack([m.forward(xs, xa) for m in ens.models])
    disc = ens.compute_discrepancy(xs, xa)
    out["ns/dims"] = np.array([S, A, N, 1] + hidden)
    for i, t in enumerate(ens.transformations):
        out[f"ns/tf{i}"] = t.numpy()
    # weights are NOT stored (42 MB): the oracle regenerates them from the seed; pin them by checksums
    out["ns/wsum"] = np.array([[float(v.double().sum()) for v in m.model.state_dict().values()] for m in ens.models])
    out["ns/wabs"] = np.array([[float(v.double().abs().sum()) for v in m.model.state_dict().values()] for m in ens.models])
    out["ns/xs"], out["ns/xa"] = xs.numpy(), xa.numpy()
    out["ns/preds"] = preds.numpy()
    out["ns/disc"] = disc.numpy()
    # threshold over the first 1024 dataset rows only (keeps the generator fast); stored with its row count
    sub = ref["datasets"].AmpDataset(s[:1024], a[:1024], s2[:1024])
    ens.train_dataloader = torch.utils.data.DataLoader(sub, batch_size=256, shuffle=True, num_workers=0)
    ens.compute_threshold()
    out["ns/threshold_rows"] = np.array(1024)
    out["ns/threshold"] = np.array(ens.threshold)

    # ---- RBFLinearCost on the north-star dims --------------------------------------------------------
    g = torch.Generator().manual_seed(2)
    es = torch.randn(256, S, generator=g)
    expert = torch.cat([es, es + 0.05 * torch.randn(256, S, generator=g)], dim=1)
    for tag, D in (("cost64", 64), ("cost512", 512)):
        cost = ref["linear_cost"].RBFLinearCost(expert, feature_dim=D, input_type="ss", bw_quantile=0.1,
                                                lambda_b=0.0025, seed=100)
        with torch.no_grad():
            nxt = xs + preds[1]
        pi = torch.cat([xs, nxt], dim=1)
        mmd = cost.fit_cost(pi)
        c = cost.get_costs(pi)
        ens.threshold = float(out["ns/threshold"])
        total, info = cost.get_bonus_costs(xs, xa, ens, next_states=nxt)
        out[f"{tag}/bw"] = np.array(cost.bw)
        out[f"{tag}/expert"] = expert.numpy() if tag == "cost64" else np.zeros(0)
        out[f"{tag}/rff_wsum"] = np.array([float(cost.rff.weight.data.double().sum()), float(cost.rff.bias.data.double().sum())])
        if D == 64:
            out[f"{tag}/rff_w"], out[f"{tag}/rff_b"] = cost.rff.weight.data.numpy(), cost.rff.bias.data.numpy()
        out[f"{tag}/phi_e"] = cost.phi_e.numpy()
        out[f"{tag}/next"] = nxt.numpy()
        out[f"{tag}/rep"] = cost.get_rep(pi).numpy()
        out[f"{tag}/mmd"] = np.array(mmd)
        out[f"{tag}/w"] = cost.w.numpy()
        out[f"{tag}/costs"] = c.numpy()
        out[f"{tag}/expert_cost"] = np.array(float(cost.get_expert_cost()))
        out[f"{tag}/total"] = total.numpy()
        for k in ("bonus", "ipm", "v_targ", "cost"):
            out[f"{tag}/info_{k}"] = info[k].numpy()

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB,", len(out), "arrays")


if __name__ == "__main__":
    main()
