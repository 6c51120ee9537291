"""Oracle parity at BASELINE.json's real sizes (parity gaps): the fused step + cost at all 40 000 rows of
configs[1]; a strided 4 096-row sample of a 1 M-row (chunked) call as configs[3] makes; the 8 x (1024 x 4) ensemble of
configs[4] at 4 096 rows; and the reference's constant-dataset-column edge case (scale 1e-8, datasets.py:35-40) under
fp16 / tf32 operands.  Tolerances as in tests/test_parity_gpu.py (1e-3 relative to max(|ref|, scale)); every check also
reports the plain relative L2 error of the output.
"""
import numpy as np
import pytest
import torch

from oracle import milo_oracle as mo
from tests import helpers as H
from tests.test_parity_gpu import REL, assert_close, collision_margin, make_engine, oracle_step

pytestmark = pytest.mark.gpu


def rel_l2(x, ref):
    x, ref = x.detach().double().cpu().reshape(-1), ref.detach().double().cpu().reshape(-1)
    return float((x - ref).norm() / ref.norm().clamp_min(1e-300))


def check_step(out, s, a, member, steps, c, oc, lam, thr, tag):
    """out = (next, disc, done, cost, ipm, bonus) rows matching the oracle on (s, a, member, steps)."""
    nxt, disc, done, cost, ipm, bonus = (o.cpu() for o in out)
    _, nxt_ref, _, done_ref, disc_ref = oracle_step(c, s, a, member, steps)
    cost_ref, info_ref = oc.get_bonus_costs(s, a, disc_ref, thr, next_states=nxt_ref.float())
    l2 = {"next_state": rel_l2(nxt, nxt_ref), "disc": rel_l2(disc, disc_ref), "cost": rel_l2(cost, cost_ref[:, 0])}
    print(f"{tag}: relative L2 {l2}")
    assert max(l2.values()) < REL, (tag, l2)                      # the plain relative-L2 bar
    assert_close(nxt, nxt_ref, c["tf"][1].mean().item(), what=f"{tag} next_state (rel L2 {l2['next_state']:.2e})")
    assert_close(disc, disc_ref, disc_ref.mean().item(), what=f"{tag} disc (rel L2 {l2['disc']:.2e})")
    cs = cost_ref.abs().max().item()
    assert_close(cost, cost_ref[:, 0], cs, what=f"{tag} cost (rel L2 {l2['cost']:.2e})")
    assert_close(-cost, -cost_ref[:, 0], cs, what=f"{tag} reward")
    assert_close(ipm, info_ref["ipm"][:, 0], cs, what=f"{tag} ipm")
    assert_close(bonus, info_ref["bonus"][:, 0], info_ref["bonus"].abs().max().item(), what=f"{tag} bonus")
    mism = done.bool().numpy() != done_ref.numpy()
    if mism.any():  # identical except within tolerance of a threshold
        assert (collision_margin(nxt_ref.numpy())[mism] < REL).all(), (tag, int(mism.sum()))
    return int(mism.sum())


def cost_oracle(c, s, a, member, steps, lam):
    oc = mo.RffCostOracle(H.ns_expert(), feature_dim=512, input_type="ss", bw_quantile=0.1, lambda_b=lam, seed=100)
    n = min(1024, s.shape[0])
    _, nxt_ref, _, _, _ = oracle_step(c, s[:n], a[:n], member[:n], steps[:n])
    oc.fit_cost(torch.cat([s[:n], nxt_ref.float()], dim=1))
    return oc


@pytest.mark.parametrize("split", [True, False])
def test_all_40000_rows_of_the_bench_batch_match_the_oracle(split):
    """BASELINE.json configs[1] at its real size: every one of the 40 000 rows, hi/lo cost operands on and off."""
    from amp_extensions_b200.engine import HumanoidTermination
    c = H.ns_case()
    E = 40000
    s = H.humanoid_like_states(E, seed=41)
    g = torch.Generator().manual_seed(42)
    a = torch.randn(E, 28, generator=g)
    member = torch.randint(0, 4, (E,), generator=g, dtype=torch.int32)
    steps = torch.randint(0, 300, (E,), generator=g, dtype=torch.int32)
    lam, thr = 0.0025, c["threshold"]
    oc = cost_oracle(c, s, a, member, steps, lam)
    eng = make_engine(c, "fp16")
    eng.set_termination(HumanoidTermination())
    eng.load_rff(oc.rff_weight, oc.rff_bias, split=True)
    eng.set_rff_split(split)
    out = eng.step_cost(s.cuda(), a.cuda(), member.cuda(), steps.clone().cuda(), oc.w.cuda(), lam, thr)
    check_step(out, s, a, member, steps, c, oc, lam, thr, f"40000 rows, split={split}")


def test_a_million_row_call_matches_the_oracle_on_a_strided_sample():
    """configs[3]'s batch: one step_cost call over 2^20 rows (16 chunks of 65 536); 4 096 rows spread over every chunk
    are compared with the oracle, and a separate call on just those rows must give bit-identical results (chunking
    and row position are invisible)."""
    from amp_extensions_b200.engine import HumanoidTermination
    c = H.ns_case()
    E, n = 1 << 20, 4096
    gd = torch.Generator(device="cuda").manual_seed(51)
    s = torch.randn(E, 226, device="cuda", generator=gd) * 0.3
    s[:, 0] = 0.7 + 0.4 * torch.rand(E, device="cuda", generator=gd)
    a = torch.randn(E, 28, device="cuda", generator=gd)
    member = torch.randint(0, 4, (E,), device="cuda", generator=gd, dtype=torch.int32)
    steps = torch.zeros(E, device="cuda", dtype=torch.int32)
    idx = (torch.arange(n) * (E // n) + (torch.arange(n) * 37) % (E // n)).cuda()   # every chunk, varying tile rows
    lam, thr = 0.0025, c["threshold"]
    ss, sa, sm = s[idx].cpu(), a[idx].cpu(), member[idx].cpu()
    oc = cost_oracle(c, ss, sa, sm, torch.zeros(n, dtype=torch.int32), lam)
    eng = make_engine(c, "fp16")
    eng.set_termination(HumanoidTermination())
    eng.load_rff(oc.rff_weight, oc.rff_bias, split=True)
    out = eng.step_cost(s, a, member, steps, oc.w.cuda(), lam, thr)
    sub = tuple(o[idx] for o in out)
    check_step(sub, ss, sa, sm, torch.zeros(n, dtype=torch.int32), c, oc, lam, thr, "1 M rows, strided sample")
    alone = eng.step_cost(s[idx].contiguous(), a[idx].contiguous(), member[idx].contiguous(),
                          torch.zeros(n, device="cuda", dtype=torch.int32), oc.w.cuda(), lam, thr)
    for x, y, name in zip(sub, alone, ("next", "disc", "done", "cost", "ipm", "bonus")):
        assert torch.equal(x, y), name
    assert int(steps.min()) == 1 and int(steps.max()) == 1


def test_the_8x1024_ensemble_matches_the_oracle_at_4096_rows():
    """configs[4]'s ensemble (8 members, hidden 1024 x 4, 28 discrepancy pairs) on 4 096 rows."""
    from amp_extensions_b200.engine import Engine, HumanoidTermination
    S, A, N, hidden, E = 226, 28, 8, [1024] * 4, 4096
    ws, bs = mo.init_ensemble(S, A, hidden, N, dense_connect=True, base_seed=100)
    tf = mo.get_transformations(*H.synth_dataset(4096, S, A, 0))
    c = dict(S=S, A=A, N=N, dense=True, hidden=hidden, act="relu", ws=ws, bs=bs, tf=tf)
    eng = Engine(S, A, N, hidden, dense_connect=True, activation="relu", transform=True, precision="fp16")
    eng.load_ensemble(ws, bs, tf)
    eng.set_termination(HumanoidTermination(horizon=300))
    s = H.humanoid_like_states(E, seed=61)
    g = torch.Generator().manual_seed(62)
    a = torch.randn(E, A, generator=g)
    member = torch.randint(0, N, (E,), generator=g, dtype=torch.int32)
    steps = torch.zeros(E, dtype=torch.int32)
    lam = 0.0025
    oc = cost_oracle(c, s, a, member, steps, lam)
    thr = float(mo.discrepancy_from_preds(mo.ensemble_forward(ws, bs, tf, s[:512], a[:512])).max())
    eng.load_rff(oc.rff_weight, oc.rff_bias, split=True)
    out = eng.step_cost(s.cuda(), a.cuda(), member.cuda(), steps.clone().cuda(), oc.w.cuda(), lam, thr)
    check_step(out, s, a, member, steps, c, oc, lam, thr, "8 x (1024 x 4), 4096 rows")


def test_constant_dataset_column_fp16_saturates_tf32_follows_the_reference_and_auto_picks_tf32():
    """datasets.py:35-40 gives a column that never varies the scale 1e-8; dynamics.py:225-227 divides by it.  A state
    1e-3 away from that constant then normalises to 1e5: the fp32 reference carries on, tf32 operands (fp32's exponent
    range) follow it within tolerance, fp16 operands saturate at 65 504 - and say so (simstep_saturation_count).
    DynamicsEnsemble therefore picks tf32 by itself when an input scale is degenerate and no precision was asked for."""
    from amp_extensions_b200 import AmpDataset, DynamicsEnsemble
    from amp_extensions_b200.engine import Engine
    S, A, N, hidden = 226, 28, 4, [512] * 4
    s_d, a_d, s2_d = H.synth_dataset(2048, S, A, 7)
    const_cols = [5, 100]
    for col in const_cols:                 # never varies in the dataset: mean = value, scale = 0 + 1e-8
        s_d[:, col] = 0.25
        s2_d[:, col] = 0.25
    tf = mo.get_transformations(s_d, a_d, s2_d)
    assert all(abs(float(tf[1][col]) - 1e-8) < 1e-12 for col in const_cols)
    ws, bs = mo.init_ensemble(S, A, hidden, N, dense_connect=True, base_seed=100)
    E = 512
    g = torch.Generator().manual_seed(8)
    s = torch.randn(E, S, generator=g)
    s[:, const_cols[0]] = 0.25 + 1e-3 * torch.randn(E, generator=g)      # drifted: normalises to ~1e5
    s[:, const_cols[1]] = 0.25                                           # not drifted: normalises to 0
    a = torch.randn(E, A, generator=g)
    member = torch.randint(0, N, (E,), generator=g, dtype=torch.int32)
    preds = mo.ensemble_forward(ws, bs, tf, s, a)
    assert torch.isfinite(preds).all()
    active = preds[member.long(), torch.arange(E)]
    nxt_ref = torch.from_numpy(mo.simenv_step(s.double().numpy(), active.numpy(), np.zeros(E, dtype=np.int64))[0])
    disc_ref = mo.discrepancy_from_preds(preds)
    res = {}
    for prec in ("tf32", "fp16"):
        eng = Engine(S, A, N, hidden, dense_connect=True, activation="relu", transform=True, precision=prec)
        eng.load_ensemble(ws, bs, tf)
        nxt, disc, _ = eng.step(s.cuda(), a.cuda(), member.cuda(), torch.zeros(E, dtype=torch.int32).cuda())
        res[prec] = (rel_l2(nxt.cpu() - s, nxt_ref - s.double()), rel_l2(disc, disc_ref), eng.saturation_count())
    print("constant-column case (delta rel L2, disc rel L2, saturated inputs):", res)
    assert res["tf32"][0] < REL and res["tf32"][1] < REL and res["tf32"][2] == 0
    big = int(((s[:, const_cols[0]] - 0.25).abs() / 1e-8 > 65504).sum())
    assert res["fp16"][2] >= big > 0                      # every out-of-range input was counted
    assert res["fp16"][0] > REL                           # and the saturation is visible in the result
    # the host mirror: no precision given -> tf32 for this dataset, fp16 for a healthy one
    ens = DynamicsEnsemble(S, A, AmpDataset(s_d, a_d, s2_d), None, num_models=N, hidden_sizes=hidden,
                           dense_connect=True, transform=True, base_seed=100)
    assert ens.operand_precision() == "tf32" and "tf32" in ens.precision_note
    out = ens.engine().step(s.cuda(), a.cuda(), member.cuda(), torch.zeros(E, dtype=torch.int32).cuda())
    # activations are ~1e4 here (an input of 1e5 went through the MLP): the bar is the relative L2 of the deltas
    assert rel_l2(out[0].cpu() - s, nxt_ref - s.double()) < REL
    healthy = DynamicsEnsemble(S, A, AmpDataset(*H.synth_dataset(2048, S, A, 7)), None, num_models=N,
                               hidden_sizes=hidden, dense_connect=True, transform=True, base_seed=100)
    assert healthy.operand_precision() == "fp16" and healthy.precision_note is None
