"""Generates tests/golden/train_golden.npz by running the REFERENCE's own DynamicsModel.train_step
(milo/milo/dynamics.py:236-250).  Build container only (needs /root/reference):

    python tests/golden/make_train_golden.py

Three optimisation steps of two tiny ensembles (SGD-Nesterov with gradient clipping; Adam) on stored batches:
losses, gradients after the first backward and the parameters after every step.  Pins
oracle/milo_oracle.py::TrainOracle (tests/test_oracle.py) and the CUDA training step (tests/test_train_gpu.py).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import load_reference, synth_dataset  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_golden.npz")


def main():
    ref = load_reference()
    out = {"torch_version": np.array(torch.__version__)}
    S, A, N, B = 20, 6, 2, 48
    s, a, s2 = synth_dataset(512, S, A, seed=0)
    ds = ref["datasets"].AmpDataset(s, a, s2)
    cases = {
        "sgd_dense": dict(hidden=[32, 24], dense=True, act="relu", optim={"optim": "sgd", "lr": 0.05, "momentum": 0.9},
                          clip=0.5),
        "adam_plain_tanh": dict(hidden=[16, 16], dense=False, act="tanh", optim={"optim": "adam", "lr": 0.01, "eps": 1e-8},
                                clip=0.0),
    }
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, 512, (3, N, B), generator=g)
    out["idx"] = idx.numpy()
    out["ds_s"], out["ds_a"], out["ds_s2"] = s.numpy(), a.numpy(), s2.numpy()
    for tag, c in cases.items():
        ens = ref["dynamics"].DynamicsEnsemble(S, A, ds, None, num_models=N, batch_size=B, hidden_sizes=c["hidden"],
                                               dense_connect=c["dense"], activation=c["act"], transform=True,
                                               optim_args=c["optim"], base_seed=100, num_workers=0)
        out[f"{tag}/dims"] = np.array([S, A, N, int(c["dense"]), B] + c["hidden"])
        out[f"{tag}/act"] = np.array(c["act"])
        out[f"{tag}/optim"] = np.array(c["optim"]["optim"])
        out[f"{tag}/lr"] = np.array(c["optim"]["lr"])
        out[f"{tag}/clip"] = np.array(c["clip"])
        for i, t in enumerate(ens.transformations):
            out[f"{tag}/tf{i}"] = t.numpy()
        for k, m in enumerate(ens.models):
            (m.state_mean, m.state_scale, m.action_mean, m.action_scale, m.diff_mean, m.diff_scale) = ens.transformations
            for name, v in m.model.state_dict().items():
                out[f"{tag}/init/m{k}/{name}"] = v.numpy().copy()
            losses = []
            for step in range(3):
                bi = idx[step, k]
                if step == 0:   # gradients of the first batch, before clipping
                    m.optimizer.zero_grad()
                    pred = m.forward(s[bi], a[bi], unnormalize_out=False)
                    target = ((s2[bi] - s[bi]) - m.diff_mean) / m.diff_scale
                    m.loss_fn(pred, target).backward()
                    for name, p in m.model.named_parameters():
                        out[f"{tag}/grad0/m{k}/{name}"] = p.grad.detach().numpy().copy()
                    out[f"{tag}/val0/m{k}"] = np.array(m.validate_step(s[bi], a[bi], s2[bi]))
                losses.append(m.train_step(c["clip"], s[bi], a[bi], s2[bi]))
                for name, v in m.model.state_dict().items():
                    out[f"{tag}/step{step}/m{k}/{name}"] = v.numpy().copy()
            out[f"{tag}/loss/m{k}"] = np.array(losses)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    main()
