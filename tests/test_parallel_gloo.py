"""N > 1 host logic on CPU: two gloo ranks exercise the env sharding and every small reduction the multi-GPU
path uses (amp_extensions_b200/parallel.py).  The data path itself has no collective."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from amp_extensions_b200 import parallel
    res = {}
    # env sharding by contiguous index ranges
    n = 1_000_003
    lo, hi = parallel.shard_range(n)
    res["shard"] = (lo, hi)
    # the union of the per-rank samples is what a single process would hold
    g = torch.Generator().manual_seed(7)
    full = torch.rand(50_000, generator=g, dtype=torch.float64) ** 3
    a, b = parallel.shard_range(full.numel())
    mine = full[a:b]
    res["max"] = float(parallel.all_reduce_max(mine.max()))
    nan_local = torch.tensor(float("nan") if rank == 1 else 1.0)
    res["max_nan"] = float(parallel.all_reduce_max(nan_local))
    res["sum"] = float(parallel.all_reduce_sum(mine.sum()))
    res["mean"] = float(parallel.global_mean(mine.sum().reshape(1), mine.numel())[0])
    for q in (0.1, 0.5, 0.999):
        res[f"q{q}"] = parallel.global_quantile(mine, q)
    # an empty shard on one rank must not break the quantile
    res["q_empty"] = parallel.global_quantile(full if rank == 0 else full[:0], 0.25)
    # rollout statistics over sharded env-steps
    E = 4001
    g2 = torch.Generator().manual_seed(11)
    cost = -torch.rand(E, generator=g2)
    ipm, bonus = cost * 0.9, cost * 0.1
    done = (torch.rand(E, generator=g2) < 0.1).to(torch.uint8)
    steps = torch.randint(1, 300, (E,), generator=g2, dtype=torch.int32)
    a, b = parallel.shard_range(E)
    res["stats"] = parallel.rollout_stats(cost[a:b], ipm[a:b], bonus[a:b], done[a:b], steps[a:b])
    # normalisation statistics of an offline dataset sharded unevenly across the ranks
    g3 = torch.Generator().manual_seed(13)
    ds_s = torch.randn(3001, 226, generator=g3) * 2 + 0.5
    ds_a = torch.randn(3001, 28, generator=g3)
    ds_n = ds_s + 0.05 * torch.randn(3001, 226, generator=g3)
    cut = slice(0, 1000) if rank == 0 else slice(1000, 3001)
    res["tf"] = parallel.global_transformations(ds_s[cut], ds_a[cut], ds_n[cut])
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_rank_results():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        yield [torch.load(os.path.join(d, f"rank{r}.pt"), weights_only=False) for r in range(world)]


def test_shards_tile_the_env_range(two_rank_results):
    (a0, b0), (a1, b1) = (r["shard"] for r in two_rank_results)
    assert a0 == 0 and b0 == a1 and b1 == 1_000_003 and abs((b0 - a0) - (b1 - a1)) <= 1


def test_reductions_equal_the_single_process_values(two_rank_results):
    g = torch.Generator().manual_seed(7)
    full = torch.rand(50_000, generator=g, dtype=torch.float64) ** 3
    for r in two_rank_results:
        assert r["max"] == float(full.max())
        assert np.isnan(r["max_nan"])  # NaN propagates like torch.max
        assert abs(r["sum"] - float(full.sum())) < 1e-9 * float(full.sum())
        assert abs(r["mean"] - float(full.mean())) < 1e-12
    assert two_rank_results[0] is not two_rank_results[1]


def test_global_quantile_matches_torch_quantile(two_rank_results):
    g = torch.Generator().manual_seed(7)
    full = torch.rand(50_000, generator=g, dtype=torch.float64) ** 3
    span = float(full.max() - full.min())
    for r in two_rank_results:
        for q in (0.1, 0.5, 0.999):
            ref = float(torch.quantile(full, q))
            assert abs(r[f"q{q}"] - ref) <= 1e-9 * span + 1e-12, (q, r[f"q{q}"], ref)
        assert abs(r["q_empty"] - float(torch.quantile(full, 0.25))) <= 1e-9 * span
    assert two_rank_results[0]["q0.5"] == two_rank_results[1]["q0.5"]  # every rank gets the same threshold


def test_rollout_stats_equal_the_single_process_values(two_rank_results):
    E = 4001
    g2 = torch.Generator().manual_seed(11)
    cost = -torch.rand(E, generator=g2)
    ipm, bonus = cost * 0.9, cost * 0.1
    done = (torch.rand(E, generator=g2) < 0.1)
    steps = torch.randint(1, 300, (E,), generator=g2, dtype=torch.int32)
    rew = (-cost).double()
    for r in two_rank_results:
        s = r["stats"]
        assert s["n"] == E and s["episodes_done"] == int(done.sum())
        assert abs(s["reward_mean"] - float(rew.mean())) < 1e-9
        assert abs(s["reward_std"] - float(rew.std(unbiased=False))) < 1e-9
        assert s["reward_max"] == float(rew.max()) and s["reward_min"] == float(rew.min())
        assert abs(s["int"] - float((-bonus).double().sum())) < 1e-6
        assert abs(s["ext"] - float((-ipm).double().sum())) < 1e-6
        assert abs(s["ep_len_mean"] - float(steps[done].double().mean())) < 1e-9


def test_sharded_transformations_equal_the_single_process_values(two_rank_results):
    """parallel.global_transformations against AmpDataset.get_transformations (itself pinned to the reference's
    datasets.py by tests/test_host_shims.py) on the union of the shards; fp32 statistics, sums carried in fp64."""
    sys.path.insert(0, ROOT)
    from amp_extensions_b200 import AmpDataset
    g3 = torch.Generator().manual_seed(13)
    ds_s = torch.randn(3001, 226, generator=g3) * 2 + 0.5
    ds_a = torch.randn(3001, 28, generator=g3)
    ds_n = ds_s + 0.05 * torch.randn(3001, 226, generator=g3)
    ref = AmpDataset(ds_s, ds_a, ds_n).get_transformations()
    for r in two_rank_results:
        for got, want in zip(r["tf"], ref):
            assert got.dtype == torch.float32 and got.shape == want.shape
            torch.testing.assert_close(got, want, rtol=2e-6, atol=1e-7)
    for x, y in zip(two_rank_results[0]["tf"], two_rank_results[1]["tf"]):
        assert torch.equal(x, y)


def test_single_process_path_needs_no_process_group():
    from amp_extensions_b200 import parallel
    assert parallel.world_size() == 1 and parallel.rank() == 0
    assert parallel.shard_range(10) == (0, 10)
    assert parallel.shard_range(10, 1, 4) == (3, 6) and parallel.shard_range(10, 3, 4) == (8, 10)
    x = torch.arange(101, dtype=torch.float64)
    assert abs(parallel.global_quantile(x, 0.2) - 20.0) < 1e-9
    assert float(parallel.all_reduce_max(torch.tensor(3.0))) == 3.0
