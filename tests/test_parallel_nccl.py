"""Two B200s, one process each, NCCL: env shards stepped on separate GPUs reproduce the single-GPU results, and
the small reductions (threshold MAX, fit_cost feature sums, rollout statistics, histogram quantile) agree with a
single-process evaluation.  Skipped on a box with fewer than two GPUs."""
import os
import socket
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E_TOTAL = 3001


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(device, rank, world):
    from amp_extensions_b200 import AmpDataset, DynamicsEnsemble, RBFLinearCost
    c = H.tiny_case("tiny_dense")
    s, a, s2 = c["ds"]
    ens = DynamicsEnsemble(c["S"], c["A"], AmpDataset(s, a, s2), None, num_models=c["N"], hidden_sizes=c["hidden"],
                           dense_connect=True, transform=True, base_seed=100, device=device)
    ens.train_dataset = AmpDataset(s[rank::world], a[rank::world], s2[rank::world])
    expert = torch.cat([s[:128], s2[:128]], dim=1)
    cost = RBFLinearCost(expert, feature_dim=64, input_type="ss", bw_quantile=0.1, lambda_b=0.1, seed=100,
                         device=device)
    g = torch.Generator().manual_seed(3)
    xs, xa = torch.randn(E_TOTAL, c["S"], generator=g), torch.randn(E_TOTAL, c["A"], generator=g)
    member = torch.randint(0, c["N"], (E_TOTAL,), generator=g, dtype=torch.int32)
    return c, ens, cost, xs, xa, member


def _run(ens, cost, xs, xa, member, device, parallel):
    eng = ens.engine()
    eng.load_rff(cost.rff.weight.data, cost.rff.bias.data, split=True)
    thr = parallel.global_threshold(ens)
    xs, xa, member = xs.to(device), xa.to(device), member.to(device)
    steps = torch.zeros(xs.shape[0], device=device, dtype=torch.int32)
    nxt, disc, done = eng.step(xs, xa, member, steps.clone())
    w = parallel.global_fit_cost(cost, torch.cat([xs, nxt], dim=1)).to(device)
    nxt, disc, done, cst, ipm, bonus = eng.step_cost(xs, xa, member, steps, w, cost.lambda_b, thr)
    stats = parallel.rollout_stats(cst, ipm, bonus, done, steps)
    q = parallel.global_quantile(disc, 0.9, engine=eng)
    return dict(thr=thr, w=w.cpu(), nxt=nxt.cpu(), disc=disc.cpu(), cost=cst.cpu(), done=done.cpu(), stats=stats, q=q)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    from amp_extensions_b200 import parallel
    c, ens, cost, xs, xa, member = _setup(device, rank, world)
    a, b = parallel.shard_range(E_TOTAL)
    res = _run(ens, cost, xs[a:b], xa[a:b], member[a:b], device, parallel)
    res["range"] = (a, b)
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_shards_match_single_gpu():
    from amp_extensions_b200 import parallel
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        parts = [torch.load(os.path.join(d, f"rank{r}.pt"), weights_only=False) for r in range(world)]
    device = torch.device("cuda", 0)
    c, ens, cost, xs, xa, member = _setup(device, 0, 1)
    ref = _run(ens, cost, xs, xa, member, device, parallel)
    assert parts[0]["range"][1] == parts[1]["range"][0] and parts[1]["range"][1] == E_TOTAL
    for p in parts:
        assert p["thr"] == ref["thr"]                       # max is exact
        assert torch.allclose(p["w"], ref["w"], atol=1e-6)   # fp64 feature sums, different order
        assert abs(p["q"] - ref["q"]) < 1e-6 * abs(ref["q"])
        for k in ("n", "episodes_done"):
            assert p["stats"][k] == ref["stats"][k]
        for k in ("reward_mean", "reward_std", "reward_max", "reward_min", "int", "ext"):
            assert abs(p["stats"][k] - ref["stats"][k]) <= 1e-5 * max(1.0, abs(ref["stats"][k])), k
    for k in ("nxt", "disc", "done"):   # rows are independent: sharding is invisible, bit for bit
        assert torch.equal(torch.cat([p[k] for p in parts]), ref[k]), k
    assert torch.allclose(torch.cat([p["cost"] for p in parts]), ref["cost"], atol=1e-5)
