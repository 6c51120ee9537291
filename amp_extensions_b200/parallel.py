"""Multi-GPU plumbing: envs shard by index, one process per GPU, full ensemble replica per GPU.

The env step itself needs no data-path collective (every env is independent).
torch.distributed (NCCL on GPUs, gloo in the CPU tests) carries only small reductions:

  * the discrepancy threshold = dataset maximum (reference milo/milo/dynamics.py:145-152) -> all-reduce MAX,
    or, as an extension, a global quantile of the per-row discrepancies via a two-pass histogram all-reduce;
  * fit_cost's feature mean (reference milo/milo/linear_cost.py:84-94) -> all-reduce SUM of [D] sums + count;
  * rollout cost / return statistics (reference mjrl/mjrl/algos/batch_reinforce.py:135-141, 288-295);
  * normalisation statistics of a sharded offline dataset (reference milo/milo/datasets.py:23-43).

Every function works un-initialised (world size 1) and on CPU tensors (gloo) as well as CUDA tensors (NCCL).
"""
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized()


def world_size():
    return dist.get_world_size() if is_dist() else 1


def rank():
    return dist.get_rank() if is_dist() else 0


def shard_range(n, rank_=None, world=None):
    """Contiguous env-index range [start, stop) owned by a rank; sizes differ by at most one."""
    r = rank() if rank_ is None else rank_
    w = world_size() if world is None else world
    base, rem = divmod(int(n), w)
    start = r * base + min(r, rem)
    return start, start + base + (1 if r < rem else 0)


def bind_host_to_gpu(device_index):
    """Pin this process to the CPU cores (and thereby the NUMA node) closest to its GPU before any pinned host
    buffer is allocated: with one process per GPU the host<->device copies of all ranks otherwise share whichever
    socket the launcher happened to start them on.  Returns the CPU list, or None when NVML / affinity is not
    available (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in cpus if c in allowed)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def _all_reduce(t, op):
    if is_dist():
        dist.all_reduce(t, op=op)
    return t


def all_reduce_max(x):
    """Global maximum of a scalar / tensor (element-wise), NaN-propagating like torch.max."""
    x = torch.as_tensor(x).clone()
    nan = torch.isnan(x).to(x.dtype)
    _all_reduce(nan, dist.ReduceOp.MAX)
    x = torch.nan_to_num(x, nan=-float("inf"))
    _all_reduce(x, dist.ReduceOp.MAX)
    return torch.where(nan > 0, torch.full_like(x, float("nan")), x)


def all_reduce_sum(x):
    return _all_reduce(torch.as_tensor(x).clone(), dist.ReduceOp.SUM)


def global_threshold(ensemble):
    """DynamicsEnsemble.compute_threshold over a dataset sharded across ranks: each rank takes the maximum
    over its own train_dataset on its GPU, one all-reduce(MAX) of a single float makes it global."""
    local = ensemble.dataset_discrepancy_max()
    ensemble.threshold = float(all_reduce_max(local).item())
    return ensemble.threshold


def global_fit_cost(cost, data_pi_local):
    """RBFLinearCost.fit_cost with the rollout rows sharded across ranks: w = global mean phi - phi_e."""
    eng = cost._precise_engine() if hasattr(cost, "_precise_engine") else cost.engine()
    n_local = int(data_pi_local.shape[0])
    if n_local > 0:
        _, psum = eng.rff_features(data_pi_local, want_sum=True)
    else:
        psum = torch.zeros(cost.feature_dim, device=eng.device, dtype=torch.float64)
    packed = torch.cat([psum, torch.tensor([float(n_local)], device=psum.device, dtype=torch.float64)])
    packed = all_reduce_sum(packed)
    phi = (packed[:-1] / packed[-1].clamp_min(1.0)).float().cpu()
    cost.w = phi - cost.phi_e
    return cost.w


def global_mean(sum_local, count_local):
    """Mean of a quantity whose per-rank sums and counts are given (fp64)."""
    packed = torch.cat([torch.as_tensor(sum_local, dtype=torch.float64).reshape(-1),
                        torch.as_tensor([float(count_local)], dtype=torch.float64).to(torch.as_tensor(sum_local).device)])
    packed = all_reduce_sum(packed)
    return packed[:-1] / packed[-1].clamp_min(1.0)


def global_transformations(states_local, actions_local, next_states_local):
    """AmpDataset.get_transformations (datasets.py:23-43) over an offline dataset whose rows are sharded across
    ranks: two passes, each one all-reduce(SUM) of the per-rank column sums (fp64) — first the means, then the mean
    absolute deviations around the GLOBAL means.  Returns the reference's tuple (state_mean, state_scale,
    action_mean, action_scale, diff_mean, diff_scale) as fp32 tensors, identical on every rank.  Host-side
    statistics, computed once when the ensemble is built (the reference does the same on the host)."""
    parts = [states_local.double(), actions_local.double(), (next_states_local - states_local).double()]
    n_local = float(states_local.shape[0])
    widths = [p.shape[1] for p in parts]
    packed = torch.cat([p.sum(dim=0) for p in parts] + [torch.tensor([n_local], dtype=torch.float64)])
    packed = all_reduce_sum(packed)
    n = packed[-1].clamp_min(1.0)
    means = [m.float() for m in torch.split(packed[:-1] / n, widths)]
    dev = torch.cat([(p.float() - m).abs().double().sum(dim=0) for p, m in zip(parts, means)])
    dev = all_reduce_sum(dev)
    scales = [(d / n).float() + 1e-8 for d in torch.split(dev, widths)]
    return (means[0], scales[0], means[1], scales[1], means[2], scales[2])


def rollout_stats(cost, ipm, bonus, done, num_steps):
    """Global rollout statistics of one batch of env-steps (batch_reinforce.py:135-141, 288-295):
    sums / extrema of reward = -cost, int = -bonus, ext = -ipm, finished episodes and their lengths."""
    f = torch.float64
    done_b = done.to(torch.bool)
    n = torch.tensor(float(cost.numel()), device=cost.device, dtype=f)
    sums = torch.stack([(-cost).to(f).sum(), ((-cost).to(f) ** 2).sum(), (-bonus).to(f).sum(), (-ipm).to(f).sum(),
                        done_b.to(f).sum(), (num_steps.to(f) * done_b.to(f)).sum(), n])
    big = float("inf")
    mx = torch.stack([(-cost).to(f).max() if cost.numel() else torch.tensor(-big, device=cost.device, dtype=f),
                      cost.to(f).max() if cost.numel() else torch.tensor(-big, device=cost.device, dtype=f)])
    sums = all_reduce_sum(sums)
    mx = all_reduce_max(mx)
    count = sums[6].clamp_min(1.0)
    mean = sums[0] / count
    var = (sums[1] / count - mean * mean).clamp_min(0.0)
    return {
        "n": int(sums[6].item()), "reward_mean": float(mean), "reward_std": float(var.sqrt()),
        "reward_max": float(mx[0]), "reward_min": float(-mx[1]), "int": float(sums[2]), "ext": float(sums[3]),
        "episodes_done": int(sums[4].item()),
        "ep_len_mean": float(sums[5] / sums[4].clamp_min(1.0)),
    }


def _global_quantile_device(x32, q, bins, refine, engine):
    """Device-resident variant: the windows around the two bracketing order statistics live in device memory
    (simstep_quantile_op), so the call issues 1 all-gather of [-min, max, count], refine + 1 all-reduces of the
    two histograms and 1 all-reduce of the window minima - and synchronises with the host ONCE, for the result."""
    dev = x32.device
    f64 = dict(device=dev, dtype=torch.float64)
    bins = min(int(bins), 4096)
    head = torch.empty(3, **f64)
    engine.quantile_op(engine.QOP_MINMAX, x32, out=head)
    head[2] = float(x32.numel())
    if is_dist():
        allh = torch.empty(world_size() * 3, **f64)
        dist.all_gather_into_tensor(allh, head)
        allh = allh.view(-1, 3)
        mm, n = allh[:, :2].max(dim=0).values, allh[:, 2].sum()
    else:
        mm, n = head[:2], head[2]
    lo, hi = -mm[0], mm[1]
    pos = q * (n - 1.0)
    k0 = torch.floor(pos).clamp_min(0.0)
    frac = (pos - k0).clamp(0.0, 1.0)
    k1 = torch.minimum(k0 + 1.0, (n - 1.0).clamp_min(0.0))
    zero = torch.zeros((), **f64)
    qstate = torch.stack([lo, hi, zero, k0, lo, hi, zero, k1]).contiguous()
    counts = torch.empty(2 * bins, device=dev, dtype=torch.int64)
    for _ in range(int(refine) + 1):
        engine.quantile_op(engine.QOP_HIST, x32, bins=bins, qstate=qstate, counts=counts)
        _all_reduce(counts, dist.ReduceOp.SUM)
        engine.quantile_op(engine.QOP_SELECT, None, bins=bins, qstate=qstate, counts=counts)
    out = torch.empty(2, **f64)
    engine.quantile_op(engine.QOP_WINMIN, x32, qstate=qstate, out=out)
    _all_reduce(out, dist.ReduceOp.MAX)
    v0, v1 = -out[0], -out[1]
    res = torch.where(n > 0, v0 + frac * (v1 - v0), torch.full((), float("nan"), **f64))
    return float(res.item())


def global_quantile(x_local, q, bins=4096, refine=2, engine=None):
    """q-quantile (linear interpolation between order statistics, as torch.quantile) of the union of every
    rank's x_local, without gathering the samples: all-reduce of min/max and of fixed-range histograms, refined
    `refine` times around the two order statistics that bracket the quantile.  Exact up to the final bin width
    (range / bins**(refine+1)); used for threshold_mode='quantile', an extension over the reference's dataset
    maximum.  With `engine` and a CUDA fp32 vector everything but the final read stays on the device
    (_global_quantile_device); CPU tensors (the gloo tests) take the torch path below."""
    xt = torch.as_tensor(x_local)
    if engine is not None and xt.is_cuda and xt.dtype == torch.float32:
        return _global_quantile_device(xt.reshape(-1).contiguous(), float(q), bins, refine, engine)
    x32 = torch.as_tensor(x_local).reshape(-1) if engine is not None else None
    x = torch.as_tensor(x_local).to(torch.float64).reshape(-1)
    dev = x.device
    n = all_reduce_sum(torch.tensor([float(x.numel())], device=dev, dtype=torch.float64))[0]
    if n.item() == 0:
        return float("nan")
    inf = float("inf")
    lo = -all_reduce_max(torch.tensor([-(x.min().item() if x.numel() else inf)], device=dev, dtype=torch.float64))[0]
    hi = all_reduce_max(torch.tensor([x.max().item() if x.numel() else -inf], device=dev, dtype=torch.float64))[0]
    pos = q * (n.item() - 1)
    k0 = int(pos)
    frac = pos - k0

    def order_stat(k):
        a, b = float(lo), float(hi)
        below = 0  # samples strictly below the current window
        for _ in range(refine + 1):
            if b <= a:
                return a
            width = (b - a) / bins
            if engine is not None and x32.is_cuda and x32.dtype == torch.float32:
                hist = engine.histogram(x32, a, b, bins).to(torch.float64)
            else:
                idx = torch.clamp(((x - a) / width).floor(), 0, bins - 1).long()
                sel = (x >= a) & (x <= b)
                hist = torch.bincount(idx[sel], minlength=bins).to(torch.float64)
            hist = all_reduce_sum(hist)
            cum = torch.cumsum(hist, 0) + below
            bin_i = int(torch.searchsorted(cum, torch.tensor([float(k) + 0.5], device=dev, dtype=torch.float64)).item())
            bin_i = min(bin_i, bins - 1)
            below = int(cum[bin_i - 1].item()) if bin_i > 0 else below
            a, b = a + bin_i * width, a + (bin_i + 1) * width
        # all samples left in the window are within one final bin width: take the window's lower edge + local min
        in_win = x[(x >= a) & (x <= b)]
        cand = torch.tensor([-(in_win.min().item() if in_win.numel() else inf)], device=dev, dtype=torch.float64)
        return float(-all_reduce_max(cand)[0])

    v0 = order_stat(k0)
    if frac == 0.0 or k0 + 1 >= int(n.item()):
        return v0
    v1 = order_stat(k0 + 1)
    return v0 + frac * (v1 - v0)
