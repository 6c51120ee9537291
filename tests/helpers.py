"""Shared fixtures: golden vectors and seeded synthetic inputs."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "milo_golden.npz")
_cache = {}


def golden():
    if "g" not in _cache:
        _cache["g"] = np.load(GOLDEN, allow_pickle=False)
    return _cache["g"]


def t(x):
    return torch.from_numpy(np.asarray(x))


def tiny_case(tag):
    """Returns dict with dims, weights/biases per member, transforms, inputs and the reference outputs."""
    g = golden()
    dims = g[f"{tag}/dims"].tolist()
    S, A, N, dense = dims[:4]
    hidden = dims[4:]
    nl = len(hidden) + 1
    ws = [[t(g[f"{tag}/m{k}/fc_layers.{l}.weight"]) for l in range(nl)] for k in range(N)]
    bs = [[t(g[f"{tag}/m{k}/fc_layers.{l}.bias"]) for l in range(nl)] for k in range(N)]
    tf = tuple(t(g[f"{tag}/tf{i}"]) for i in range(6))
    return dict(S=S, A=A, N=N, dense=bool(dense), hidden=hidden, act=str(g[f"{tag}/act"]), ws=ws, bs=bs, tf=tf,
                xs=t(g[f"{tag}/xs"]), xa=t(g[f"{tag}/xa"]), preds=t(g[f"{tag}/preds"]),
                preds_norm=t(g[f"{tag}/preds_norm"]), disc=t(g[f"{tag}/disc"]), threshold=float(g[f"{tag}/threshold"]),
                ds=(t(g[f"{tag}/ds_s"]), t(g[f"{tag}/ds_a"]), t(g[f"{tag}/ds_s2"])))


def synth_dataset(M, S, A, seed):
    """Same generator as tests/golden/make_golden.py::synth_dataset."""
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(M, S, generator=g)
    a = torch.randn(M, A, generator=g)
    s2 = s + 0.05 * torch.randn(M, S, generator=g)
    return s, a, s2


def ns_case():
    """North-star ensemble: weights regenerated from the seed by the oracle, pinned by stored checksums."""
    if "ns" in _cache:
        return _cache["ns"]
    from oracle import milo_oracle as mo
    g = golden()
    S, A, N = 226, 28, 4
    hidden = [512] * 4
    ws, bs = mo.init_ensemble(S, A, hidden, N, dense_connect=True, base_seed=100)
    tf = tuple(t(g[f"ns/tf{i}"]) for i in range(6))
    case = dict(S=S, A=A, N=N, dense=True, hidden=hidden, act="relu", ws=ws, bs=bs, tf=tf, xs=t(g["ns/xs"]),
                xa=t(g["ns/xa"]), preds=t(g["ns/preds"]), disc=t(g["ns/disc"]), threshold=float(g["ns/threshold"]),
                threshold_rows=int(g["ns/threshold_rows"]), wsum=g["ns/wsum"], wabs=g["ns/wabs"])
    _cache["ns"] = case
    return case


def ns_expert():
    g = torch.Generator().manual_seed(2)
    es = torch.randn(256, 226, generator=g)
    return torch.cat([es, es + 0.05 * torch.randn(256, 226, generator=g)], dim=1)


def humanoid_like_states(E, seed, fall_fraction=0.3):
    """States in the 226-d humanoid3d layout whose collision test has both outcomes: root height in
    s[0], body positions (relative to root) at 9b+1.., rotation normal at 9b+4.., velocities from 136."""
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(E, 226, generator=g) * 0.3
    s[:, 0] = 0.7 + 0.4 * torch.rand(E, generator=g)
    for b in range(15):
        s[:, 9 * b + 2] = (torch.rand(E, generator=g) - 0.5) * 0.8  # relative y of body b
        s[:, 9 * b + 5] = torch.rand(E, generator=g) * 2 - 1        # normal.y
    low = torch.rand(E, generator=g) < fall_fraction
    s[low, 0] = 0.05 + 0.3 * torch.rand(int(low.sum()), generator=g)
    s[:, 136:] = torch.randn(E, 90, generator=g) * 3.0
    return s


def spinkick_raw():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "amp_extensions_b200", "data",
                             "humanoid3d_spinkick.npz"))
    return z["frames_raw"]


def perturbed_poses(E, seed, clip=None, t_max=None, with_origin=True, ch=None):
    """SURVEY.md section 8d generator: pose = clip(t) perturbed (root pos += N(0,.05), every quaternion
    <- exp(N(0,.2) axis) q, revolute += N(0,.2)), vel = clipvel(t) + N(0,.5), t ~ U(0, t_max)."""
    from oracle import imitation_oracle as io
    ch = ch or io.HUMANOID3D
    clip = clip or io.Clip(spinkick_raw(), ch, "wrap")
    rng = np.random.default_rng(seed)
    offs, sizes = io.param_layout(ch)
    dof = int(sum(sizes))
    t = rng.uniform(0, t_max if t_max is not None else clip.duration, E)
    origin = rng.normal(0, 0.3, (E, 3)) if with_origin else None
    pose, vel = np.zeros((E, dof)), np.zeros((E, dof))
    for e in range(E):
        p = clip.kin_pose(t[e], origin[e] if with_origin else (0, 0, 0))
        if with_origin:
            p[1] -= origin[e, 1]  # the simulated character stands on the real ground
        v = clip.kin_vel(t[e]) + rng.normal(0, 0.5, dof)
        p[0:3] += rng.normal(0, 0.05, 3)
        for j, jt in enumerate(ch["joint_type"]):
            o = offs[j] + (3 if jt == io.ROOT else 0)
            if jt in (io.ROOT, io.SPHERICAL):
                ax = rng.normal(0, 1, 3)
                ang = rng.normal(0, 0.2)
                dq = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax / np.linalg.norm(ax)])
                q = io.quat_mul(dq, p[o:o + 4])
                p[o:o + 4] = q / np.linalg.norm(q)
                v[o + 3] = 0.0  # 4th slot of an angular velocity is unused (KinTree.cpp:1546)
            elif jt == io.REVOLUTE:
                p[o] += rng.normal(0, 0.2)
        pose[e], vel[e] = p, v
    return pose, vel, t, origin


def trainstep_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trainstep_golden.npz"))


def trainstep_paths(g):
    """The rollout of tests/golden/trainstep_golden.npz as the list of path dicts train_step works on."""
    paths = []
    for k in range(len(g["length"])):
        n = int(g["length"][k])
        paths.append(dict(observations=g["observations"][k, :n].copy(), next_observations=g["next_observations"][k, :n].copy(),
                          actions=g["actions"][k, :n].copy(), rewards=np.zeros(n)))
    return paths


def reward_replacement(paths, reward_func, ensemble):
    """BatchREINFORCE.train_step's reward replacement (mjrl/mjrl/algos/batch_reinforce.py:103-169, cost input 'ss',
    no GAIL), written against the objects' public methods only, so the same lines run over the reference classes,
    over the oracle and over this package's shims.  Mutates paths[i]['rewards']; returns the infos dict."""
    infos = {"int": [], "ext": [], "reward": [], "ep_len": []}
    cost_input = np.concatenate([np.concatenate([t["observations"], t["next_observations"]], axis=1) for t in paths], axis=0)
    infos["mb_mmd"] = reward_func.fit_cost(torch.from_numpy(cost_input).float())                    # BR:113
    for traj in paths:
        states = torch.from_numpy(traj["observations"]).float()
        next_states = torch.from_numpy(traj["next_observations"]).float()
        actions = torch.from_numpy(traj["actions"]).float()
        bonus_cost, cost_info = reward_func.get_bonus_costs(states, actions, ensemble, next_states=next_states)  # BR:128
        bonus_cost = bonus_cost[:, 0]
        intrinsic_sum = -np.sum(cost_info["bonus"][:, 0].numpy())                                   # BR:135-136
        extrinsic_sum = -np.sum(cost_info["ipm"][:, 0].numpy())
        infos["int"].append(intrinsic_sum)
        infos["ext"].append(extrinsic_sum)
        infos["reward"].append(extrinsic_sum + intrinsic_sum)
        infos["ep_len"].append(len(traj["rewards"]))
        traj["rewards"] = -1.0 * bonus_cost.cpu().numpy()                                           # BR:144
    infos["bonus_mmd"] = float(np.concatenate([-1.0 * t["rewards"] for t in paths], axis=0).mean()
                               - float(reward_func.get_expert_cost()))                               # BR:169
    return infos


def character_dict_from_json(d):
    """A DeepMimic character file (parsed JSON) as the table oracle/imitation_oracle.py works on.  Body attach
    rotations are dropped: they turn the body's own frame, not its centre of mass (KinTree.cpp:1156-1166), so the
    imitation reward does not see them."""
    from oracle import imitation_oracle as io
    types = {"none": io.ROOT, "spherical": io.SPHERICAL, "revolute": io.REVOLUTE, "fixed": io.FIXED}
    joints = d["Skeleton"]["Joints"]
    bodies = {b["ID"]: b for b in d["BodyDefs"]}
    attach = [(j["AttachX"], j["AttachY"], j["AttachZ"]) for j in joints]
    attach[0] = (0.0, 0.0, 0.0)   # KinTree.cpp:1064-1067
    return dict(joint_type=[types[j["Type"]] for j in joints], parent=[j["Parent"] for j in joints], attach=attach,
                is_end_eff=[int(j.get("IsEndEffector", 0)) for j in joints],
                diff_weight=[float(j.get("DiffWeight", 1.0)) for j in joints],
                body_mass=[float(bodies[j["ID"]]["Mass"]) for j in joints],
                body_attach=[(bodies[j["ID"]]["AttachX"], bodies[j["ID"]]["AttachY"], bodies[j["ID"]]["AttachZ"])
                             for j in joints])
