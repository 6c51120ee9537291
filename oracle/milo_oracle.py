"""CPU oracle for the MILO part of the hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in torch-CPU fp32 (the reference's own arithmetic) and numpy float64 (SimEnv's state),
of the reference functions on the path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product (amp_extensions_b200) never
does and fails loudly without its CUDA library.

Pinning: tests/golden/milo_golden.npz was produced by tests/golden/make_golden.py, which imports the
reference's own milo/milo/{dynamics,datasets,linear_cost}.py in the build container and stores their
outputs; tests/test_oracle.py checks every function below against those vectors.  SimEnv imports gym and
the SWIG DeepMimicCore, which are absent; tests/golden/make_simenv_golden.py stubs exactly those two imports
and runs the reference's gym_simenv/envs/sim_env.py unmodified (constructor on the reference's own arg /
character / controller files, reset, step, is_done, check_*), so `simenv_*` is pinned by
tests/golden/simenv_golden.npz (episodes, 116 contact-threshold states, velocity check) as well as by the
hand-built known-answer states.

Reference files (paths under the reference tree):
  DYN = milo/milo/dynamics.py     DS = milo/milo/datasets.py
  LC  = milo/milo/linear_cost.py  SE = gym-simenv/gym_simenv/envs/sim_env.py
"""
import math

import numpy as np
import torch
import torch.nn as nn

# --------------------------------------------------------------------------------------
# datasets.py


def get_transformations(states, actions, next_states):
    """AmpDataset.get_transformations (DS:23-43): mean and mean-absolute-deviation (+1e-8)."""
    diff = next_states - states
    state_mean = states.mean(dim=0).float()
    action_mean = actions.mean(dim=0).float()
    diff_mean = diff.mean(dim=0).float()
    state_scale = torch.abs(states - state_mean).mean(dim=0).float() + 1e-8
    action_scale = torch.abs(actions - action_mean).mean(dim=0).float() + 1e-8
    diff_scale = torch.abs(diff - diff_mean).mean(dim=0).float() + 1e-8
    return state_mean, state_scale, action_mean, action_scale, diff_mean, diff_scale


# --------------------------------------------------------------------------------------
# dynamics.py


def layer_input_sizes(input_dim, output_dim, hidden_sizes, dense_connect):
    """Fan-in of every nn.Linear of BasicMLP (DYN:412-420)."""
    sizes = [input_dim] + list(hidden_sizes) + [output_dim]
    fan_in = []
    for i in range(len(sizes) - 1):
        k = sizes[i]
        if dense_connect:
            k += sum(sizes[:i])
        fan_in.append(k)
    return fan_in, sizes[1:]


def init_member(state_dim, action_dim, hidden_sizes, dense_connect, seed):
    """Random-init weights exactly as DynamicsModel.__init__ does (DYN:184-196): seed torch and
    numpy, then construct the nn.Linear layers in order.  Returns ([W_l], [b_l])."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    fan_in, fan_out = layer_input_sizes(state_dim + action_dim, state_dim, hidden_sizes, dense_connect)
    ws, bs = [], []
    for k, o in zip(fan_in, fan_out):
        lin = nn.Linear(k, o)
        ws.append(lin.weight.detach().clone())
        bs.append(lin.bias.detach().clone())
    return ws, bs


def init_ensemble(state_dim, action_dim, hidden_sizes, num_models, dense_connect=True, base_seed=100):
    """DynamicsEnsemble.__init__ (DYN:70-80): member k is seeded with base_seed + k."""
    members = [init_member(state_dim, action_dim, hidden_sizes, dense_connect, base_seed + k)
               for k in range(num_models)]
    return [m[0] for m in members], [m[1] for m in members]


def mlp_forward(ws, bs, x, dense_connect=True, activation="relu"):
    """BasicMLP.forward (DYN:422-433)."""
    act = torch.relu if activation == "relu" else torch.tanh
    inp = x
    for w, b in zip(ws[:-1], bs[:-1]):
        out = act(torch.nn.functional.linear(inp, w, b))
        inp = torch.cat([inp, out], dim=1) if dense_connect else out
    return torch.nn.functional.linear(inp, ws[-1], bs[-1])


def dynamics_forward(ws, bs, transforms, state, action, dense_connect=True, activation="relu",
                     unnormalize_out=True):
    """DynamicsModel.forward (DYN:216-233).  transforms = the 6-tuple of get_transformations or None."""
    state = state.float()
    action = action.float()
    if transforms is not None:
        state_mean, state_scale, action_mean, action_scale, diff_mean, diff_scale = transforms
        state = (state - state_mean) / state_scale
        action = (action - action_mean) / action_scale
    diff = mlp_forward(ws, bs, torch.cat([state, action], dim=1), dense_connect, activation)
    if transforms is not None and unnormalize_out:
        diff = diff * diff_scale + diff_mean
    return diff


def ensemble_forward(all_ws, all_bs, transforms, state, action, dense_connect=True, activation="relu"):
    """Stack of every member's un-normalised prediction, [N, B, S] (first line of DYN:139)."""
    with torch.no_grad():
        return torch.stack([dynamics_forward(ws, bs, transforms, state, action, dense_connect, activation)
                            for ws, bs in zip(all_ws, all_bs)], dim=0)


def discrepancy_from_preds(preds):
    """Pairwise-max L2 discrepancy of DynamicsEnsemble.compute_discrepancy (DYN:140-143)."""
    n = preds.shape[0]
    if n < 2:
        return torch.zeros(preds.shape[1])
    disc = torch.stack([torch.norm(preds[i] - preds[j], p=2, dim=1) for i in range(n) for j in range(i + 1, n)], dim=0)
    return disc.max(0).values


def compute_discrepancy(all_ws, all_bs, transforms, state, action, dense_connect=True, activation="relu"):
    """DynamicsEnsemble.compute_discrepancy (DYN:134-143)."""
    return discrepancy_from_preds(ensemble_forward(all_ws, all_bs, transforms, state, action, dense_connect, activation))


def compute_threshold(all_ws, all_bs, transforms, states, actions, batch_size=256, dense_connect=True):
    """DynamicsEnsemble.compute_threshold (DYN:145-152): dataset maximum of the discrepancy (the
    DataLoader shuffling of the reference does not change a maximum)."""
    best = -math.inf
    for i in range(0, states.shape[0], batch_size):
        d = compute_discrepancy(all_ws, all_bs, transforms, states[i:i + batch_size].float(),
                                actions[i:i + batch_size].float(), dense_connect)
        best = max(best, d.max().item())
    return best


# --------------------------------------------------------------------------------------
# linear_cost.py


class RffCostOracle:
    """RBFLinearCost (LC:6-152) restated; same constructor arguments, same RNG consumption order."""

    def __init__(self, expert_data, feature_dim=1024, input_type="ss", cost_range=(-1.0, 0.0), bw_quantile=0.1,
                 bw_samples=100000, lambda_b=1.0, seed=100):
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.expert_data = expert_data
        self.input_type = input_type
        self.feature_dim = feature_dim
        self.cost_range = cost_range
        if cost_range is not None:
            self.c_min, self.c_max = cost_range
        self.lambda_b = lambda_b
        input_dim = expert_data.size(1)
        # fit_bandwidth (LC:73-82)
        n = expert_data.shape[0]
        i0 = torch.randint(low=0, high=n, size=(bw_samples,))
        i1 = torch.randint(low=0, high=n, size=(bw_samples,))
        norm = torch.norm(expert_data[i0, :] - expert_data[i1, :], dim=1)
        self.bw = torch.quantile(norm, q=bw_quantile).item()
        # rff layer (LC:53-55)
        lin = nn.Linear(input_dim, feature_dim)
        self.rff_bias = ((torch.rand_like(lin.bias.data) - 0.5) * 2.0 * np.pi).detach()
        self.rff_weight = (torch.rand_like(lin.weight.data) / (self.bw + 1e-8)).detach()
        self.w = None
        self.expert_rep = self.get_rep(expert_data)
        self.phi_e = self.expert_rep.mean(dim=0)

    def get_rep(self, x):
        """LC:64-71."""
        with torch.no_grad():
            out = torch.nn.functional.linear(x.cpu().float(), self.rff_weight, self.rff_bias)
            return torch.cos(out) * np.sqrt(2 / self.feature_dim)

    def fit_cost(self, data_pi):
        """LC:84-94."""
        phi = self.get_rep(data_pi).mean(0)
        self.w = phi - self.phi_e
        return torch.dot(self.w, self.w).item()

    def get_costs(self, x):
        """LC:96-103."""
        data = self.get_rep(x)
        out = torch.mm(data, self.w.unsqueeze(1))
        if self.cost_range is not None:
            return torch.clamp(out, self.c_min, self.c_max)
        return out

    def get_expert_cost(self):
        """LC:105-109."""
        return (1 - self.lambda_b) * torch.clamp(torch.mm(self.expert_rep, self.w.unsqueeze(1)), self.c_min,
                                                 self.c_max).mean()

    def rff_input(self, states, actions, next_states):
        """LC:115-126."""
        if self.input_type == "sa":
            return torch.cat([states, actions], dim=1)
        if self.input_type == "ss":
            return torch.cat([states, next_states], dim=1)
        if self.input_type == "sas":
            return torch.cat([states, actions, next_states], dim=1)
        if self.input_type == "s":
            return states
        raise NotImplementedError("Input type not implemented")

    def get_bonus_costs(self, states, actions, discrepancy, threshold, next_states=None):
        """LC:111-152 with the ensemble call (LC:132) replaced by its result `discrepancy`."""
        rff_cost = self.get_costs(self.rff_input(states, actions, next_states))
        if self.cost_range is not None:
            d = (discrepancy / threshold).view(-1, 1).clone()
            d[d > 1.0] = 1.0
            bonus = d * self.c_min
        else:
            bonus = discrepancy.view(-1, 1)
        ipm = (1 - self.lambda_b) * rff_cost
        weighted_bonus = self.lambda_b * bonus
        cost = ipm - weighted_bonus
        return cost, {"bonus": weighted_bonus, "ipm": ipm, "v_targ": rff_cost, "cost": cost}


class MlpCostOracle(RffCostOracle):
    """MLPCost (LC:154-301) restated: the feature map is an MLP ending in tanh, then cos(.) * sqrt(2/D).  Same
    constructor arguments and RNG consumption order as the reference (net first, then the bandwidth draw)."""

    def __init__(self, expert_data, hidden_dims=(2048, 2048), activation="relu", feature_dim=1024, input_type="ss",
                 cost_range=(-1.0, 0.0), bw_quantile=0.1, bw_samples=100000, lambda_b=1.0, seed=100):
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.expert_data = expert_data
        self.input_type = input_type
        self.feature_dim = feature_dim
        self.cost_range = cost_range
        if cost_range is not None:
            self.c_min, self.c_max = cost_range
        self.lambda_b = lambda_b
        input_dim = expert_data.size(1)
        hidden_dims = list(hidden_dims)
        # LC:201-214 (including its quirk: with a single hidden size the net is Linear(in, hidden[0]) + Tanh)
        self.act = torch.relu if activation == "relu" else torch.tanh
        dim = feature_dim if not hidden_dims else hidden_dims[0]
        linears = [nn.Linear(input_dim, dim)]
        if hidden_dims[1:]:
            for size in hidden_dims[1:]:
                linears.append(nn.Linear(dim, size))
                dim = size
            linears.append(nn.Linear(dim, feature_dim))
        self.ws = [l.weight.data.clone() for l in linears]
        self.bs = [l.bias.data.clone() for l in linears]
        # fit_bandwidth (LC:223-229): unused by the features but it advances the RNG and is reported
        n = expert_data.shape[0]
        i0 = torch.randint(low=0, high=n, size=(bw_samples,))
        i1 = torch.randint(low=0, high=n, size=(bw_samples,))
        self.bw = torch.quantile(torch.norm(expert_data[i0, :] - expert_data[i1, :], dim=1), q=bw_quantile).item()
        self.w = None
        self.expert_rep = self.get_rep(expert_data)
        self.phi_e = self.expert_rep.mean(0)

    def get_rep(self, x):
        """LC:231-236."""
        with torch.no_grad():
            out = x.cpu().float()
            for i in range(len(self.ws) - 1):
                out = self.act(torch.nn.functional.linear(out, self.ws[i], self.bs[i]))
            out = torch.tanh(torch.nn.functional.linear(out, self.ws[-1], self.bs[-1]))
            return torch.cos(out) * np.sqrt(2 / self.feature_dim)


def gail_disc_forward(ws, bs, x, activation="relu"):
    """Discriminator.forward (milo/milo/gail_cost.py:18-43): Linear / activation stack, last layer linear."""
    act = torch.relu if activation == "relu" else torch.tanh
    out = x.cpu().float()
    with torch.no_grad():
        for i in range(len(ws) - 1):
            out = act(torch.nn.functional.linear(out, ws[i], bs[i]))
        return torch.nn.functional.linear(out, ws[-1], bs[-1])


def gail_costs(disc_outs, loss_type="least_squares"):
    """GAILCost.get_ls_costs / get_ll_costs (gail_cost.py:232-253)."""
    if loss_type == "least_squares":
        rewards = 1.0 - 0.25 * (1.0 - disc_outs) ** 2
        rewards[rewards < 0.0] = 0.0
        return -rewards
    return torch.nn.functional.logsigmoid(disc_outs)


def gail_bonus_costs(input_cost, discrepancy, lambda_b):
    """GAILCost.get_bonus_costs (gail_cost.py:255-283) after the input concatenation and the discriminator."""
    ipm = (1 - lambda_b) * input_cost
    bonus = lambda_b * discrepancy.view(-1, 1)
    cost = ipm - bonus
    return cost, {"bonus": bonus, "ipm": ipm, "v_targ": input_cost, "cost": cost}


class TrainOracle:
    """DynamicsModel.train_step / validate_step for one member (DYN:236-262) restated with torch autograd on the
    CPU: BasicMLP forward on normalised inputs (DYN:422-433), MSE against the normalised difference, backward,
    optional clip_grad_norm_, torch.optim.SGD(nesterov=True) or torch.optim.Adam (DYN:198-203)."""

    def __init__(self, ws, bs, transforms, dense_connect=True, activation="relu", optim_args=None):
        optim_args = optim_args or {"optim": "sgd", "lr": 1e-4, "momentum": 0.9}
        self.ws = [nn.Parameter(w.clone().float()) for w in ws]
        self.bs = [nn.Parameter(b.clone().float()) for b in bs]
        self.tf = transforms
        self.dense, self.activation = dense_connect, activation
        params = [p for pair in zip(self.ws, self.bs) for p in pair]   # nn.Module order: weight, bias per layer
        self.params = params
        if optim_args["optim"] == "sgd":
            self.opt = torch.optim.SGD(params, lr=optim_args["lr"], momentum=optim_args["momentum"], nesterov=True)
        else:
            self.opt = torch.optim.Adam(params, lr=optim_args["lr"], eps=optim_args["eps"])

    def _loss(self, state, action, next_state):
        sm, ss, am, as_, dm, dsc = self.tf if self.tf is not None else (0, 1, 0, 1, 0, 1)
        x = torch.cat([(state - sm) / ss, (action - am) / as_], dim=-1) if self.tf is not None else \
            torch.cat([state, action], dim=-1)
        act = torch.relu if self.activation == "relu" else torch.tanh
        inp = x
        for i in range(len(self.ws) - 1):
            out = act(torch.nn.functional.linear(inp, self.ws[i], self.bs[i]))
            inp = torch.cat([inp, out], dim=-1) if self.dense else out
        pred = torch.nn.functional.linear(inp, self.ws[-1], self.bs[-1])
        target = next_state - state
        if self.tf is not None:
            target = (target - dm) / dsc
        return torch.nn.functional.mse_loss(pred, target)

    def validate_step(self, state, action, next_state):
        with torch.no_grad():
            return self._loss(state, action, next_state).item()

    def grads(self, state, action, next_state):
        self.opt.zero_grad()
        loss = self._loss(state, action, next_state)
        loss.backward()
        return loss.item()

    def train_step(self, grad_clip, state, action, next_state):
        loss = self.grads(state, action, next_state)
        if grad_clip:
            nn.utils.clip_grad_norm_(self.params, grad_clip)
        self.opt.step()
        return loss


# --------------------------------------------------------------------------------------
# sim_env.py  (state in float64, as the reference keeps it)

HUMANOID3D_FALL_BODIES = (0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 12, 13, 14)  # SE:102
# humanoid3d.txt BodyDefs: shape, Param0 (diameter), Param1 (height)
HUMANOID3D_BODY_DEFS = {
    0: ("sphere", 0.18, 0.18), 1: ("sphere", 0.22, 0.22), 2: ("sphere", 0.205, 0.205),
    3: ("capsule", 0.11, 0.30), 4: ("capsule", 0.10, 0.31), 5: ("box", 0.177, 0.055),
    6: ("capsule", 0.09, 0.18), 7: ("capsule", 0.08, 0.135), 8: ("sphere", 0.08, 0.08),
    9: ("capsule", 0.11, 0.30), 10: ("capsule", 0.10, 0.31), 11: ("box", 0.177, 0.055),
    12: ("capsule", 0.09, 0.18), 13: ("capsule", 0.08, 0.135), 14: ("sphere", 0.08, 0.08),
}


def simenv_collided(ob, bodies=HUMANOID3D_FALL_BODIES, body_defs=HUMANOID3D_BODY_DEFS, pos_dim=3, rot_dim=6,
                    record_all_world=False, record_world_root_pos=False):
    """SimEnv.check_collision / check_sphere / check_capsule (SE:175-257) on a batch ob[B, S]."""
    ob = np.asarray(ob, dtype=np.float64)
    collided = np.zeros(ob.shape[0], dtype=bool)
    for index, body in enumerate(bodies):
        shape, p0, p1 = body_defs[body]
        offset = (pos_dim + rot_dim) * body + 1
        if record_all_world or (index == 0 and record_world_root_pos):
            y = ob[:, offset + 1]
        else:
            y = ob[:, 0] + ob[:, offset + 1]
        radius = 0.5 * p0
        if shape == "sphere":
            collided |= y <= radius + 0.0001
        elif shape == "capsule":
            norm_y = ob[:, offset + pos_dim + 1]
            top = y + 0.5 * p1 * norm_y
            bottom = y - 0.5 * p1 * norm_y
            collided |= (top <= radius + 0.0001) | (bottom <= radius + 0.0001)
    return collided


def simenv_velocity_exploded(ob, vel_offset=136, threshold=100.0, divisor=1.0):
    """SimEnv.check_velocity (SE:259-268), without the reference's in-place scaling of ob."""
    v = np.asarray(ob, dtype=np.float64)[:, vel_offset:] / divisor
    return np.any(np.abs(v) > threshold, axis=1)


def simenv_step(ob, delta_active, num_steps, horizon=300, enable_velocity_check=False, **collision_kw):
    """SimEnv.step + is_done (SE:140-173) for a batch: ob f64[B,S] += f32 delta; returns
    (next_ob f64, num_steps+1, done)."""
    nxt = np.asarray(ob, dtype=np.float64) + np.asarray(delta_active, dtype=np.float32).astype(np.float64)
    steps = np.asarray(num_steps) + 1
    done = (steps >= horizon) | simenv_collided(nxt, **collision_kw)
    if enable_velocity_check:
        done |= simenv_velocity_exploded(nxt)
    return nxt, steps, done
