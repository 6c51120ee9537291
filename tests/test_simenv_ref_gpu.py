"""GPU: the env plugin and the step kernel against the reference's own sim_env.py.

tests/golden/simenv_golden.npz was written by tests/golden/make_simenv_golden.py, which runs the reference SimEnv
class unmodified (only `gym` and the SWIG simulator are stubbed).  Here the CUDA path replays the same episodes
through `amp_extensions_b200.SimEnv` (reset / step / done / member round-robin) and the same hand-placed contact and
velocity states through the C ABI's step.

Tolerance: states within 1e-3 of the state scale (1.0) per step of a free-running episode (fp16 operands, fp32
accumulate; the error compounds over the ten steps, so the bound is 1e-3 * steps); termination flags identical
except where a tested height sits within 1e-3 of its threshold in the reference trajectory.
"""
import os

import numpy as np
import pytest
import torch

from oracle import milo_oracle as mo
from tests import helpers as H
from tests.test_parity_gpu import collision_margin

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "simenv_golden.npz"))
S, A = 226, 28


def build_ensemble():
    from amp_extensions_b200 import AmpDataset, DynamicsEnsemble
    N, hidden = int(GOLD["N"]), [int(h) for h in GOLD["hidden"]]
    ds = AmpDataset(*H.synth_dataset(int(GOLD["dataset_rows"]), S, A, int(GOLD["dataset_seed"])))
    ens = DynamicsEnsemble(S, A, ds, None, num_models=N, hidden_sizes=hidden, dense_connect=True, activation="relu",
                           transform=True, base_seed=int(GOLD["base_seed"]))
    got = [[float(l.weight.data.double().abs().sum()) for l in m.model.fc_layers] for m in ens.models]
    np.testing.assert_allclose(got, GOLD["weight_checksum"], rtol=1e-12)   # the reference's init, from the seed
    return ens


def test_plugin_episodes_match_the_reference_simenv():
    from amp_extensions_b200 import SimEnv
    ens = build_ensemble()
    queue = [GOLD["traj_ob0"][e] for e in range(GOLD["traj_ob0"].shape[0])]
    env = SimEnv(ens, horizon=8, seed=7, reset_fn=lambda n, rng: queue.pop(0)[None, :])
    assert env.dynamics is ens.models[0]
    worst = 0.0
    for e in range(GOLD["traj_ob0"].shape[0]):
        ob = env.reset()
        assert ob.dtype == np.float64 and np.array_equal(ob, GOLD["traj_ob0"][e])
        assert env.reset_counter == int(GOLD["traj_member"][e]) and env.dynamics is ens.models[env.reset_counter]
        for k in range(GOLD["traj_actions"].shape[1]):
            ob, reward, done, info = env.step(GOLD["traj_actions"][e, k].copy())
            ref = GOLD["traj_obs"][e, k]
            err = float(np.abs(ob - ref).max())
            worst = max(worst, err / (k + 1))
            assert err < 1e-3 * (k + 1), (e, k, err)
            assert reward == 0 and env.num_steps == int(GOLD["traj_num_steps"][e, k])
            if bool(done) != bool(GOLD["traj_done"][e, k]):
                assert collision_margin(ref[None])[0] < 1e-3, (e, k)
    assert GOLD["traj_done"][:, 0].any() and not GOLD["traj_done"][:, 0].all()   # early falls and horizon cuts both occur
    print(f"worst per-step state error {worst:.2e}")


def zero_delta_engine(enable_velocity_check, **term_kw):
    """An ensemble that predicts delta = 0, so the step's termination test sees exactly the state it was given."""
    from amp_extensions_b200.engine import Engine, HumanoidTermination
    hidden = [32, 32]
    ws, bs = mo.init_ensemble(S, A, hidden, 2, base_seed=5, dense_connect=True)
    ws = [[torch.zeros_like(w) for w in m] for m in ws]
    bs = [[torch.zeros_like(b) for b in m] for m in bs]
    tf = (torch.zeros(S), torch.ones(S), torch.zeros(A), torch.ones(A), torch.zeros(S), torch.ones(S))
    eng = Engine(S, A, 2, hidden, dense_connect=True, activation="relu", transform=True, precision="fp16")
    eng.load_ensemble(ws, bs, tf)
    eng.set_termination(HumanoidTermination(horizon=300, enable_velocity_check=enable_velocity_check, **term_kw))
    return eng


def run_done(eng, states):
    E = states.shape[0]
    s = torch.from_numpy(states).float().cuda()
    nxt, _, done = eng.step(s, torch.zeros(E, A, device="cuda"), torch.zeros(E, dtype=torch.int32, device="cuda"),
                            torch.zeros(E, dtype=torch.int32, device="cuda"))
    assert torch.equal(nxt, s)
    return done.cpu().numpy().astype(bool)


def test_contact_thresholds_match_the_reference_simenv():
    """Every fall body, both capsule caps, 2e-6 and 1e-3 either side of `<= radius + 1e-4` (sim_env.py:188, 236)."""
    done = run_done(zero_delta_engine(False), GOLD["contact_states"])
    assert (done == GOLD["contact_collided"]).all(), np.nonzero(done != GOLD["contact_collided"])


def test_velocity_check_matches_the_reference_simenv():
    assert (run_done(zero_delta_engine(True), GOLD["vel_states"]) == GOLD["vel_done_enabled"]).all()
    assert (run_done(zero_delta_engine(False), GOLD["vel_states"]) == GOLD["vel_done_default"]).all()


def test_controller_flags_match_the_reference_simenv():
    """RecordAllWorld / RecordWorldRootPos / RecordVelAsPos as the reference SimEnv honours them."""
    for k, (aw, wrp) in enumerate(GOLD["variant_flag_sets"]):
        eng = zero_delta_engine(False, record_all_world=bool(aw), record_world_root_pos=bool(wrp))
        done = run_done(eng, GOLD["variant_states"][k])
        assert (done == GOLD["variant_collided"][k]).all(), (aw, wrp, np.nonzero(done != GOLD["variant_collided"][k]))
    eng = zero_delta_engine(True, vel_divisor=float(GOLD["velpos_divisor"]))
    assert (run_done(eng, GOLD["velpos_states"]) == GOLD["velpos_done"]).all()
