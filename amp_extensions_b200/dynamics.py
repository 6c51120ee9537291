"""Dynamics ensemble with the reference's Python surface, evaluated on B200 through libsimstep.

Mirrors reference milo/milo/dynamics.py:19-233 (DynamicsEnsemble, DynamicsModel) and the state-dict
layout of BasicMLP (dynamics.py:394-420), so `ensemble.pt` files written by either side load in the
other (dynamics.py:110-131), SimEnv can index `.models[i]` and call `.forward`/`.model.eval()`
(sim_env.py:119-120, 157, 282-284) and RBFLinearCost can call `.get_action_discrepancy` and read
`.threshold` (linear_cost.py:132).

What is different by design: every member is evaluated in ONE grouped tensor-core launch per layer,
so `models[i].forward` and `compute_discrepancy` share a single pass.  `DynamicsEnsemble.train` trains all members
at once on the device (each on its own shuffled batches, dynamics.py:82-108, 236-250, 264-380); a reference-trained
object can also be wrapped with `DynamicsEnsemble.from_reference`.
"""
import numpy as np
import torch
import torch.nn as nn

from . import engine as _engine


class MLPParams(nn.Module):
    """Parameter container with BasicMLP's module layout: fc_layers.{i}.{weight,bias}, fan-in of layer i =
    layer_sizes[i] + sum(layer_sizes[:i]) when dense_connect (dynamics.py:412-420).  It deliberately has no
    forward: evaluation is DynamicsModel.forward on the GPU."""

    def __init__(self, input_dim, output_dim, hidden_sizes, dense_connect=False, activation="relu"):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.dense_connect = dense_connect
        self.activation = activation
        self.layer_sizes = [input_dim] + list(hidden_sizes) + [output_dim]
        layers = []
        for i in range(len(self.layer_sizes) - 1):
            fan_in = self.layer_sizes[i]
            if dense_connect:
                fan_in += sum(self.layer_sizes[:i])
            layers.append(nn.Linear(fan_in, self.layer_sizes[i + 1]))
        self.fc_layers = nn.ModuleList(layers)


class DynamicsModel:
    """One ensemble member (dynamics.py:167-233). Predicts the state DIFFERENCE."""

    def __init__(self, state_dim, action_dim, hidden_sizes=(512, 512), use_resnet=False, dense_connect=True,
                 activation="relu", transform=True, optim_args=None, device=torch.device("cpu"), seed=100):
        if use_resnet:
            raise NotImplementedError("ResidualMLP members are outside the accelerated path; use the reference class")
        torch.manual_seed(seed)  # dynamics.py:184-186: the member's init is a function of its seed
        np.random.seed(seed)
        self.state_dim = state_dim
        self.action_dim = action_dim
        self.transform = transform
        self.device = device
        self.model = MLPParams(state_dim + action_dim, state_dim, list(hidden_sizes), dense_connect, activation)
        self._optim_state = None
        self._ensemble = None
        self._index = None
        self.state_mean = self.state_scale = self.action_mean = None
        self.action_scale = self.diff_mean = self.diff_scale = None

    def forward(self, state, action, unnormalize_out=True):
        """dynamics.py:216-233. Returns a CUDA tensor [B, state_dim]."""
        if self._ensemble is None:
            raise RuntimeError("DynamicsModel.forward needs the owning DynamicsEnsemble (members share one launch)")
        if isinstance(state, np.ndarray):
            state = torch.from_numpy(state).float()
        if isinstance(action, np.ndarray):
            action = torch.from_numpy(action).float()
        diff = self._ensemble.forward_all(state, action)[self._index]
        if self.transform and not unnormalize_out:
            dev = diff.device
            diff = (diff - self.diff_mean.to(dev)) / self.diff_scale.to(dev)
        return diff

    def load(self, model_state_dict, optimizer_state_dict=None):
        """dynamics.py:380-386."""
        self.model.load_state_dict(model_state_dict)
        if optimizer_state_dict:
            self._optim_state = optimizer_state_dict
        if self._ensemble is not None:
            self._ensemble.mark_dirty()

    def get_state_dicts(self):
        """dynamics.py:388-392."""
        return {"model": self.model.state_dict(), "optim": self._optim_state if self._optim_state else {}}

    def train(self, *args, **kwargs):
        raise NotImplementedError("members are trained together: call DynamicsEnsemble.train() (all members step in "
                                  "one grouped pass on the device)")


class DynamicsEnsemble:
    """dynamics.py:19-165 with the same constructor signature."""

    def __init__(self, state_dim, action_dim, train_dataset, validate_dataset, num_models=4, batch_size=256,
                 hidden_sizes=(512, 512), use_resnet=False, dense_connect=True, activation="relu", transform=True,
                 optim_args=None, device=torch.device("cpu"), base_seed=100, num_workers=1, precision=None,
                 max_chunk_envs=0):
        self.state_dim = state_dim
        self.action_dim = action_dim
        self.train_dataset = train_dataset
        self.validate_dataset = validate_dataset
        self.batch_size = batch_size
        self.num_models = num_models
        self.hidden_sizes = list(hidden_sizes)
        self.dense_connect = dense_connect
        self.activation = activation
        self.transform = transform
        self.device = torch.device(device)
        self.base_seed = base_seed
        self.precision = precision
        self.max_chunk_envs = max_chunk_envs
        self.optim_args = dict(optim_args) if optim_args else {"optim": "sgd", "lr": 1e-4, "momentum": 0.9}
        self.transformations = None
        if transform:
            if train_dataset is None:
                raise ValueError("transform=True needs a train_dataset providing get_transformations()")
            self.transformations = tuple(t.detach().cpu() for t in train_dataset.get_transformations(torch.device("cpu")))
        self.models = [DynamicsModel(state_dim, action_dim, hidden_sizes=self.hidden_sizes, use_resnet=use_resnet,
                                     dense_connect=dense_connect, activation=activation, transform=transform,
                                     optim_args=optim_args, device=device, seed=base_seed + k)
                       for k in range(num_models)]
        for k, m in enumerate(self.models):
            m._ensemble, m._index = self, k
        self._assign_transforms()
        self.threshold = 0.0
        self._eng = None
        self._param_stamp = None

    # -- construction helpers -----------------------------------------------------------
    @classmethod
    def from_reference(cls, ref, precision=None, device=None):
        """Wrap a (trained or loaded) reference milo.dynamics.DynamicsEnsemble: copies member state dicts,
        transformations and threshold."""
        first = ref.models[0].model
        sizes = list(first.layer_sizes)
        self = cls.__new__(cls)
        self.state_dim, self.action_dim = ref.state_dim, ref.action_dim
        self.train_dataset = getattr(ref.train_dataloader, "dataset", None) if hasattr(ref, "train_dataloader") else None
        self.validate_dataset = None
        self.batch_size = 256
        self.num_models = len(ref.models)
        self.hidden_sizes = sizes[1:-1]
        self.dense_connect = bool(first.dense_connect)
        self.activation = "relu" if first.nonlinearity is torch.relu else "tanh"
        self.transform = bool(ref.transform)
        self.device = torch.device(device if device is not None else "cuda")
        self.base_seed = getattr(ref, "base_seed", 100)
        self.precision = precision
        self.max_chunk_envs = 0
        self.optim_args = {"optim": "sgd", "lr": 1e-4, "momentum": 0.9}
        self.transformations = tuple(t.detach().cpu() for t in ref.transformations) if ref.transform else None
        self.models = []
        for k, rm in enumerate(ref.models):
            m = DynamicsModel(self.state_dim, self.action_dim, hidden_sizes=self.hidden_sizes,
                              dense_connect=self.dense_connect, activation=self.activation, transform=self.transform,
                              seed=self.base_seed + k)
            m.model.load_state_dict({k2: v.detach().cpu() for k2, v in rm.model.state_dict().items()})
            m._ensemble, m._index = self, k
            self.models.append(m)
        self._assign_transforms()
        self.threshold = float(getattr(ref, "threshold", 0.0))
        self._eng = None
        self._param_stamp = None
        return self

    # -- training on the device ------------------------------------------------------------
    def train(self, epochs, validate=False, logger=None, log_epoch=False, grad_clip=0, save_path=None,
              save_checkpoints=False, writer=None, seed=None, graph=True):
        """dynamics.py:82-108 over DynamicsModel.train (:264-380), with all members stepping together: per epoch
        every member walks its own random permutation of the train dataset in batches of `batch_size`
        (DataLoader(shuffle=True), dynamics.py:58); train_step = forward on normalised inputs, MSE, backward,
        optional clip_grad_norm_, SGD-Nesterov or Adam (dynamics.py:236-250, 198-203) on a tf32 training handle.
        After training each member holds the parameters of its best-training-loss epoch (dynamics.py:370-372).
        Returns [(train_min_loss, first_epoch_loss)] per member like the reference."""
        if validate:
            raise NotImplementedError("DynamicsEnsemble.train(validate=True): the validation pass and its best_validate "
                                      "checkpoint (dynamics.py:82-108, 252-268) stay with the reference class")
        if save_checkpoints:
            raise NotImplementedError("DynamicsEnsemble.train(save_checkpoints=True): per-model epoch checkpoints "
                                      "(dynamics.py:355-378) stay with the reference class; save_ensemble() writes "
                                      "the reference's ensemble.pt")
        ds = self.train_dataset
        n = len(ds)
        B = min(int(self.batch_size), n)
        oa = self.optim_args
        dev = self.device if self.device.type == "cuda" else None
        eng = _engine.Engine(self.state_dim, self.action_dim, self.num_models, self.hidden_sizes,
                             dense_connect=self.dense_connect, activation=self.activation, transform=self.transform,
                             precision="tf32", device=dev)
        eng.train_init(B, optim=oa.get("optim", "sgd"), lr=oa.get("lr", 1e-4), momentum=oa.get("momentum", 0.9),
                       eps=oa.get("eps", 1e-8))
        ws = [[l.weight.data for l in m.model.fc_layers] for m in self.models]
        bs = [[l.bias.data for l in m.model.fc_layers] for m in self.models]
        eng.load_ensemble(ws, bs, self.transformations)
        d = eng.device
        S_all = ds.states.to(d, torch.float32)
        A_all = ds.actions.to(d, torch.float32)
        N_all = ds.next_states.to(d, torch.float32)
        gen = torch.Generator(device=d)
        gen.manual_seed(int(self.base_seed if seed is None else seed))
        N = self.num_models
        best = [float("inf")] * N
        best_params = [None] * N
        first = [None] * N
        history = []
        for epoch in range(int(epochs)):
            perms = torch.stack([torch.randperm(n, device=d, generator=gen) for _ in range(N)])
            tot = torch.zeros(N, device=d, dtype=torch.float64)
            nb = 0
            for i0 in range(0, n, B):
                idx = perms[:, i0:i0 + B]                       # the last batch may be short (drop_last=False)
                step = eng.train_step_graph if graph else eng.train_step   # full batches replay one CUDA graph
                tot += step(S_all[idx], A_all[idx], N_all[idx], grad_clip=float(grad_clip or 0.0))
                nb += 1
            avg = (tot / nb).cpu().tolist()                      # np.average of the batch losses, dynamics.py:283
            history.append(avg)
            improved = [k for k in range(N) if avg[k] < best[k]]
            if improved:
                pw, pb = eng.train_export(eng.TRAIN_PARAMS)
                for k in improved:
                    best[k] = avg[k]
                    best_params[k] = (pw[k], pb[k])
            for k in range(N):
                if first[k] is None:
                    first[k] = avg[k]
            if logger is not None and log_epoch:
                logger.info("Epoch: {}, Train Loss: {}".format(epoch, avg))
            if writer is not None:
                for k in range(N):
                    writer.add_scalar(f"Loss/train/model{k}", avg[k], epoch)
        for k, m in enumerate(self.models):                      # dynamics.py:370-372: keep the best-train-loss epoch
            if best_params[k] is not None:
                for l, layer in enumerate(m.model.fc_layers):
                    layer.weight.data = best_params[k][0][l].clone()
                    layer.bias.data = best_params[k][1][l].clone()
        self.mark_dirty()
        self.train_history = history
        eng.close()
        if save_path is not None:
            self.save_ensemble(save_path if str(save_path).endswith(".pt") else str(save_path) + "/ensemble.pt")
        return [(best[k], first[k]) for k in range(N)]

    def _assign_transforms(self):
        if self.transform and self.transformations is not None:
            for m in self.models:  # dynamics.py:128-131
                (m.state_mean, m.state_scale, m.action_mean, m.action_scale, m.diff_mean,
                 m.diff_scale) = self.transformations

    # -- device side --------------------------------------------------------------------
    def mark_dirty(self):
        self._param_stamp = None

    def _stamp(self):
        return tuple((id(p), p._version) for m in self.models for p in m.model.parameters())

    def engine(self):
        """The libsimstep handle with the current weights packed (re-packed when parameters changed)."""
        if self._eng is None:
            dev = self.device if self.device.type == "cuda" else None
            self._eng = _engine.Engine(self.state_dim, self.action_dim, self.num_models, self.hidden_sizes,
                                       dense_connect=self.dense_connect, activation=self.activation,
                                       transform=self.transform, precision=self.operand_precision(), device=dev,
                                       max_chunk_envs=self.max_chunk_envs)
            self._param_stamp = None
        stamp = self._stamp()
        if stamp != self._param_stamp:
            ws = [[l.weight.data for l in m.model.fc_layers] for m in self.models]
            bs = [[l.bias.data for l in m.model.fc_layers] for m in self.models]
            self._eng.load_ensemble(ws, bs, self.transformations)
            self._param_stamp = stamp
        return self._eng

    # A dataset column that never varies gets scale 1e-8 (datasets.py:35-40); the reference then feeds its fp32 MLP
    # (s - mean) / 1e-8, which leaves the fp16 range (65 504) as soon as a state is 1e-3 away from that constant.
    DEGENERATE_SCALE = 1e-6

    def operand_precision(self):
        """The tensor-core operand format of this ensemble's handle: the caller's choice when one was given;
        otherwise fp16 (twice the tf32 rate) unless the normalisation has a degenerate input scale, in which case
        tf32 (fp32's exponent range) is used and the choice is recorded in `self.precision_note`."""
        if self.precision is not None:
            return self.precision
        self.precision_note = None
        if self.transform and self.transformations is not None:
            scales = torch.cat([torch.as_tensor(self.transformations[1]).reshape(-1).float(),
                                torch.as_tensor(self.transformations[3]).reshape(-1).float()])
            if bool((scales < self.DEGENERATE_SCALE).any()):
                self.precision_note = (f"{int((scales < self.DEGENERATE_SCALE).sum())} input scale(s) below "
                                       f"{self.DEGENERATE_SCALE:g}: tf32 operands instead of fp16")
                return "tf32"
        return _engine.DEFAULT_PRECISION

    def forward_all(self, state, action):
        """Every member's un-normalised prediction, CUDA [N, B, S]."""
        return self.engine().forward(state, action)

    # -- reference API ------------------------------------------------------------------
    def save_ensemble(self, save_path):
        """dynamics.py:110-116: list of {'model': state_dict, 'optim': state_dict}."""
        torch.save([m.get_state_dicts() for m in self.models], save_path)

    def load_ensemble(self, state_dict_path):
        """dynamics.py:118-131."""
        state_dicts = torch.load(state_dict_path, map_location="cpu")
        assert len(state_dicts) == len(self.models)
        for model, sd in zip(self.models, state_dicts):
            model.load(sd["model"], sd.get("optim"))
        self._assign_transforms()
        self.mark_dirty()

    def compute_discrepancy(self, state, action):
        """dynamics.py:134-143. Returns a CPU tensor [B] like the reference."""
        return self.engine().discrepancy(state, action).cpu()

    def get_action_discrepancy(self, state, action):
        """dynamics.py:154-165."""
        if isinstance(state, np.ndarray):
            state = torch.from_numpy(state)
        if isinstance(action, np.ndarray):
            action = torch.from_numpy(action)
        if state.dim() == 1:
            state = state.unsqueeze(0)
        if action.dim() == 1:
            action = action.unsqueeze(0)
        return self.compute_discrepancy(state.float(), action.float())

    def dataset_discrepancy_max(self, batch_rows=1 << 16):
        """Maximum discrepancy over this process's train_dataset, as a CUDA fp64 scalar tensor."""
        ds = self.train_dataset
        eng = self.engine()
        best = torch.full((), -float("inf"), device=eng.device, dtype=torch.float64)
        n = len(ds)
        for i in range(0, n, batch_rows):
            d = eng.discrepancy(ds.states[i:i + batch_rows].float(), ds.actions[i:i + batch_rows].float())
            best = torch.maximum(best, eng.reduce_max_sum(d)[0])
        return best

    def compute_threshold(self):
        """dynamics.py:145-152: dataset maximum of the discrepancy (order independent, so no shuffling)."""
        self.threshold = float(self.dataset_discrepancy_max().item())

    # -- pickling (SimEnv is shipped to worker processes, sim_env.py:48, sampler.py:116-121) ----------
    def __getstate__(self):
        d = dict(self.__dict__)
        d["_eng"] = None
        d["_param_stamp"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        for k, m in enumerate(self.models):
            m._ensemble, m._index = self, k
