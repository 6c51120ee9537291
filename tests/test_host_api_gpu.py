"""Host-buffer entry points (amp_extensions_b200/host_api.py) against direct engine calls on the same inputs:
chunking, stream overlap and several batches in flight must be invisible in the results (bit for bit — every row
is computed by the same kernels, rows are independent)."""
import pytest
import torch

from tests import helpers as H
from tests.test_parity_gpu import make_engine

pytestmark = pytest.mark.gpu


def _setup(E):
    from amp_extensions_b200 import RBFLinearCost
    from amp_extensions_b200.engine import HumanoidTermination
    c = H.tiny_case("tiny_dense")
    eng = make_engine(c, "fp16")
    eng.set_termination(HumanoidTermination(horizon=5, fall_contact_bodies=()))
    s, a, s2 = c["ds"]
    cost = RBFLinearCost(torch.cat([s[:128], s2[:128]], dim=1), feature_dim=64, input_type="ss", bw_quantile=0.1,
                         lambda_b=0.1, seed=100)
    eng.load_rff(cost.rff.weight.data, cost.rff.bias.data, split=True)
    g = torch.Generator().manual_seed(4)
    w = (torch.randn(64, generator=g) * 0.05).cuda()
    states = [torch.randn(E, c["S"], generator=g).pin_memory() for _ in range(3)]
    actions = [torch.randn(E, c["A"], generator=g).pin_memory() for _ in range(6)]
    member = torch.randint(0, c["N"], (E,), generator=g, dtype=torch.int32).pin_memory()
    return c, eng, w, states, actions, member


def _direct(eng, s, a, member, steps, w):
    nxt, disc, done, cst, _, _ = eng.step_cost(s.cuda(), a.cuda(), member.cuda(), steps, w, 0.1, 0.7)
    return nxt, cst, done, disc


@pytest.mark.parametrize("E", [1, 300, 1025])
def test_host_step_pipeline_equals_direct_calls(E):
    from amp_extensions_b200.host_api import HostStepPipeline
    c, eng, w, states, actions, member = _setup(E)
    pipe = HostStepPipeline(eng, E, n_chunks=3, with_cost=True, depth=2)
    steps_h = torch.zeros(E, dtype=torch.int32).pin_memory()
    # two batches in flight, collected in order
    pipe.submit(states[0], actions[0], member, steps_h, w, 0.1, 0.7)
    pipe.submit(states[1], actions[1], member, steps_h, w, 0.1, 0.7)
    with pytest.raises(RuntimeError):
        pipe.submit(states[2], actions[2], member, steps_h, w, 0.1, 0.7)
    for k in range(2):
        nxt_h, cost_h, done_h, disc_h, st_h = pipe.collect()
        nxt, cst, done, disc = _direct(eng, states[k], actions[k], member, torch.zeros(E, dtype=torch.int32).cuda(), w)
        assert torch.equal(nxt_h, nxt.cpu()) and torch.equal(cost_h, cst.cpu())
        assert torch.equal(done_h, done.cpu()) and torch.equal(disc_h, disc.cpu())
        assert (st_h == 1).all()
    out = pipe.step(states[2], actions[2], member, steps_h, w, 0.1, 0.7)   # synchronous form
    nxt, cst, done, disc = _direct(eng, states[2], actions[2], member, torch.zeros(E, dtype=torch.int32).cuda(), w)
    assert torch.equal(out[0], nxt.cpu()) and torch.equal(out[1], cst.cpu())


@pytest.mark.parametrize("E", [2, 513])
def test_host_env_pipeline_keeps_state_on_the_device(E):
    """env.step(actions) with resident state: six steps of two alternating groups equal six chained direct steps,
    including the step counter reaching the horizon (done flips at step 5)."""
    from amp_extensions_b200.host_api import HostEnvPipeline
    c, eng, w, states, actions, member = _setup(E)
    pipe = HostEnvPipeline(eng, E, groups=2, n_chunks=2, with_cost=True)
    pipe.reset(0, states[0], member)
    pipe.reset(1, states[1], member)
    ref_state = [states[0].cuda(), states[1].cuda()]
    ref_steps = [torch.zeros(E, dtype=torch.int32).cuda() for _ in range(2)]
    pipe.submit(0, actions[0], w, 0.1, 0.7)
    for i in range(1, 7):
        if i < 6:
            pipe.submit(i % 2, actions[i], w, 0.1, 0.7)
        k = (i - 1) % 2
        obs_h, cost_h, done_h, disc_h, st_h = pipe.collect()
        nxt, disc, done, cst, _, _ = eng.step_cost(ref_state[k], actions[i - 1].cuda(), member.cuda(), ref_steps[k], w,
                                                   0.1, 0.7)
        ref_state[k] = nxt
        assert torch.equal(obs_h, nxt.cpu()), i
        assert torch.equal(cost_h, cst.cpu()) and torch.equal(disc_h, disc.cpu()) and torch.equal(done_h, done.cpu())
        assert torch.equal(st_h, ref_steps[k].cpu())
    with pytest.raises(RuntimeError):
        pipe.submit(0, actions[0], w, 0.1, 0.7)
        pipe.submit(0, actions[0], w, 0.1, 0.7)


@pytest.mark.parametrize("mode", ["lean", "obs_fp16"])
def test_host_env_pipeline_optional_blocks_and_half_precision_observations(mode):
    """want_disc / want_steps = False drop those blocks from the download (one copy of the record's core); with
    obs_dtype=float16 the observations come back as the fp16 rounding of the fp32 state, which itself stays exact
    on the device (the next step starts from the fp32 state, not from the rounded copy)."""
    from amp_extensions_b200.host_api import HostEnvPipeline
    E = 513
    c, eng, w, states, actions, member = _setup(E)
    kw = dict(want_disc=False, want_steps=False)
    if mode == "obs_fp16":
        kw["obs_dtype"] = torch.float16
    pipe = HostEnvPipeline(eng, E, groups=1, n_chunks=2, with_cost=True, **kw)
    pipe.reset(0, states[0], member)
    ref_state = states[0].cuda()
    ref_steps = torch.zeros(E, dtype=torch.int32).cuda()
    for i in range(3):
        pipe.submit(0, actions[i], w, 0.1, 0.7)
        obs_h, cost_h, done_h, disc_h, st_h = pipe.collect()
        nxt, disc, done, cst, _, _ = eng.step_cost(ref_state, actions[i].cuda(), member.cuda(), ref_steps, w, 0.1, 0.7)
        ref_state = nxt
        assert disc_h is None and st_h is None
        assert torch.equal(cost_h, cst.cpu()) and torch.equal(done_h, done.cpu())
        if mode == "obs_fp16":
            assert obs_h.dtype == torch.float16 and torch.equal(obs_h, nxt.cpu().to(torch.float16)), i
        else:
            assert torch.equal(obs_h, nxt.cpu()), i
    assert pipe.d2h_bytes_per_step == E * (eng.S * (2 if mode == "obs_fp16" else 4) + 1 + 4)
