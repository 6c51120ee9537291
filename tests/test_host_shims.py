"""Host-side mirror of the reference's plugin surface, without a GPU: the ensemble.pt save format
(reference milo/milo/dynamics.py:110-131), constructor / attribute compatibility, SimEnv's reset protocol and
pickling (gym-simenv/gym_simenv/envs/sim_env.py:48, 118-132, 270-285), termination tables (sim_env.py:100-116)."""
import os
import pickle

import numpy as np
import pytest
import torch

from tests import helpers as H
from amp_extensions_b200 import AmpDataset, DynamicsEnsemble, HumanoidTermination, SimEnv
from amp_extensions_b200 import _lib

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_PT = os.path.join(GOLDEN_DIR, "ref_ensemble_tiny_dense.pt")


def tiny_ensemble():
    c = H.tiny_case("tiny_dense")
    ds = AmpDataset(*c["ds"])
    ens = DynamicsEnsemble(c["S"], c["A"], ds, None, num_models=c["N"], hidden_sizes=c["hidden"],
                           dense_connect=True, activation="relu", transform=True, base_seed=100)
    return c, ens


def test_constructor_reproduces_the_reference_init_and_transforms():
    """Same seeds, same layer construction order => bit-identical random init (dynamics.py:79, 184-196)."""
    c, ens = tiny_ensemble()
    for k, m in enumerate(ens.models):
        for l, lin in enumerate(m.model.fc_layers):
            assert torch.equal(lin.weight.data, c["ws"][k][l])
            assert torch.equal(lin.bias.data, c["bs"][k][l])
    for mine, ref in zip(ens.transformations, c["tf"]):
        assert torch.equal(mine, ref)
    m0 = ens.models[0]
    assert torch.equal(m0.state_mean, c["tf"][0]) and torch.equal(m0.diff_scale, c["tf"][5])  # dynamics.py:128-131


def test_loads_an_ensemble_pt_written_by_the_reference():
    c, ens = tiny_ensemble()
    for m in ens.models:  # scramble, then load the reference's file
        for p in m.model.parameters():
            p.data.zero_()
    ens.load_ensemble(REF_PT)
    for k, m in enumerate(ens.models):
        for l, lin in enumerate(m.model.fc_layers):
            assert torch.equal(lin.weight.data, c["ws"][k][l]) and torch.equal(lin.bias.data, c["bs"][k][l])
    # optimizer state is carried through untouched so a later save keeps it (dynamics.py:385-392)
    ref_sd = torch.load(REF_PT, map_location="cpu")
    assert ens.models[0].get_state_dicts()["optim"].keys() == ref_sd[0]["optim"].keys()


def test_save_ensemble_writes_the_reference_format(tmp_path):
    c, ens = tiny_ensemble()
    path = str(tmp_path / "ensemble.pt")
    ens.save_ensemble(path)
    mine = torch.load(path, map_location="cpu")
    ref = torch.load(REF_PT, map_location="cpu")
    assert isinstance(mine, list) and len(mine) == len(ref) == c["N"]
    for a, b in zip(mine, ref):
        assert set(a.keys()) == set(b.keys()) == {"model", "optim"}
        assert list(a["model"].keys()) == list(b["model"].keys())
        for k in a["model"]:
            assert a["model"][k].dtype == b["model"][k].dtype and torch.equal(a["model"][k], b["model"][k])
    with pytest.raises(AssertionError):  # dynamics.py:123
        small = DynamicsEnsemble(c["S"], c["A"], AmpDataset(*c["ds"]), None, num_models=2, hidden_sizes=c["hidden"])
        small.load_ensemble(path)


def test_ensemble_pickles_without_its_device_handle():
    _, ens = tiny_ensemble()
    ens.threshold = 1.25
    clone = pickle.loads(pickle.dumps(ens))
    assert clone._eng is None and clone.threshold == 1.25 and len(clone.models) == len(ens.models)
    assert all(m._ensemble is clone and m._index == k for k, m in enumerate(clone.models))
    assert torch.equal(clone.models[1].model.fc_layers[0].weight, ens.models[1].model.fc_layers[0].weight)


def test_training_needs_the_device_and_members_train_together():
    """DynamicsEnsemble.train runs on the GPU only (no CPU fallback); a single member cannot be trained alone
    (the grouped pass steps all of them); ResidualMLP members are outside the accelerated path."""
    from amp_extensions_b200 import _lib
    _, ens = tiny_ensemble()
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SimstepError):
            ens.train(1)
    with pytest.raises(NotImplementedError):
        ens.models[0].train(1, None)
    with pytest.raises(NotImplementedError):
        DynamicsEnsemble(4, 2, None, None, transform=False, use_resnet=True)


def test_simenv_reset_protocol_and_member_round_robin():
    """sim_env.py:118-119: first episode = member 0; every reset advances (c+1) % N (sim_env.py:282-283);
    observations are float64 copies (sim_env.py:285)."""
    c, ens = tiny_ensemble()
    pool = np.arange(5 * c["S"], dtype=np.float32).reshape(5, c["S"])
    env = SimEnv(ens, reset_states=pool, horizon=7, seed=3)
    assert env.horizon == 7 and env.state_size == c["S"] and env.action_size == c["A"]
    assert env.observation_space.shape == (c["S"],) and env.action_space.shape == (c["A"],)
    assert env.dynamics is ens.models[0] and env.reset_counter == 0
    seen = []
    for i in range(5):
        ob = env.reset()
        assert ob.dtype == np.float64 and ob.shape == (c["S"],) and env.num_steps == 0
        assert any(np.array_equal(ob, pool[j].astype(np.float64)) for j in range(5))
        ob[0] = -1.0  # a copy: mutating it must not touch the env's state
        assert env.get_observation()[0] != -1.0
        seen.append(env.reset_counter)
        assert env.dynamics is ens.models[env.reset_counter]
    assert seen == [1, 2, 0, 1, 2]
    assert env.seed_env(11) == 11
    a = env.np_random.uniform()
    env.seed_env(11)
    assert env.np_random.uniform() == a
    with pytest.raises(AssertionError):  # sim_env.py:152
        fresh = SimEnv(ens, reset_states=pool)
        fresh.step(np.zeros(c["A"]))


def test_simenv_is_picklable_for_worker_pools():
    """sampler.py:116-121 ships the env (and through it the ensemble) to Pool workers."""
    c, ens = tiny_ensemble()
    pool = np.zeros((2, c["S"]), dtype=np.float32)
    env = SimEnv(ens, reset_states=pool, horizon=9, enable_velocity_check=True, seed=5)
    clone = pickle.loads(pickle.dumps(env))
    assert clone.horizon == 9 and clone.enable_velocity_check and clone._vec is None
    assert clone.dynamic_ensemble._eng is None
    assert clone.reset().shape == (c["S"],)
    with pytest.raises(RuntimeError):
        SimEnv(ens).reset()  # no simulator, no reset_fn, no reset_states


def test_termination_tables_follow_sim_env():
    """sim_env.py:102-104: offsets (pos_dim+rot_dim)*body + 1 over the fall-contact bodies; shapes and
    diameters / heights from humanoid3d.txt BodyDefs."""
    t = HumanoidTermination(horizon=300, enable_velocity_check=True).to_struct()
    bodies = [0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 12, 13, 14]
    assert t.n_bodies == 13 and t.horizon == 300 and t.enable_velocity_check == 1 and t.vel_offset == 136
    assert list(t.body_offset[:13]) == [9 * b + 1 for b in bodies]
    shapes = [t.body_shape[i] for i in range(13)]
    assert shapes == [0, 0, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0]
    assert t.body_param0[0] == pytest.approx(0.18) and t.body_param0[3] == pytest.approx(0.11)
    assert t.body_param1[3] == pytest.approx(0.30) and t.body_param1[6] == pytest.approx(0.135)
    assert abs(t.vel_threshold - 100.0) < 1e-6 and t.pos_dim == 3
    # dict-style BodyDefs as parsed from the character file work as well
    defs = [dict(Shape="sphere", Param0=0.2, Param1=0.2), dict(Shape="box", Param0=1, Param1=1)]
    t2 = HumanoidTermination(body_defs=defs, fall_contact_bodies=[0, 1]).to_struct()
    assert t2.n_bodies == 2 and t2.body_shape[1] == _lib.SHAPE["box"] and t2.body_offset[1] == 10


def test_amp_dataset_transformations_match_the_oracle():
    from oracle import milo_oracle as mo
    s, a, s2 = H.synth_dataset(300, 11, 3, 4)
    mine = AmpDataset(s, a, s2).get_transformations(torch.device("cpu"))
    for x, y in zip(mine, mo.get_transformations(s, a, s2)):
        assert torch.equal(x, y)


def test_simenv_constructor_reads_the_reference_scene_files(monkeypatch):
    """The `deepmimic_args=` path of the plugin (sim_env.py:50-116, 270-285): the reference's own arg / controller /
    character files parsed through the reference's ArgParser, the simulator replaced by the stand-in the golden
    generator uses.  Needs the reference tree (build container); skipped elsewhere."""
    import importlib.util
    import sys
    import types
    ref = os.environ.get("SIMSTEP_REFERENCE", "/root/reference")
    dm_root = os.path.join(ref, "deepmimic", "deepmimic")
    arg_file = "args/run_amp_humanoid3d_spinkick_args.txt"
    if not os.path.exists(os.path.join(dm_root, arg_file)):
        pytest.skip("reference tree not present")
    sys.path.insert(0, GOLDEN_DIR)
    try:
        import make_simenv_golden as msg
    finally:
        sys.path.remove(GOLDEN_DIR)
    from oracle import imitation_oracle as io
    msg.FakeSimulator.clip = io.Clip(H.spinkick_raw(), io.HUMANOID3D, "wrap")
    spec = importlib.util.spec_from_file_location("deepmimic.util.arg_parser", os.path.join(dm_root, "util", "arg_parser.py"))
    ap = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ap)
    mods = {name: types.ModuleType(name) for name in ("deepmimic", "deepmimic.env", "deepmimic.env.deepmimic_env",
                                                      "deepmimic.util")}
    mods["deepmimic.env.deepmimic_env"].DeepMimicEnv = msg.FakeSimulator
    mods["deepmimic.util.arg_parser"] = ap
    for name, m in mods.items():
        m.__path__ = [] if name in ("deepmimic", "deepmimic.env", "deepmimic.util") else getattr(m, "__path__", None)
        monkeypatch.setitem(sys.modules, name, m)
    monkeypatch.chdir(dm_root)   # the arg file names its character / controller files relative to this directory

    c, _ = tiny_ensemble()
    s, a, s2 = H.synth_dataset(64, 226, 28, 0)
    ens = DynamicsEnsemble(226, 28, AmpDataset(s, a, s2), None, num_models=2, hidden_sizes=[16], dense_connect=True,
                           transform=True, base_seed=100)
    env = SimEnv(ens, deepmimic_args=arg_file, horizon=11, seed=7, enable_velocity_check=True)
    assert env.deepmimic is not None and env.time_max == pytest.approx(msg.FakeSimulator.clip.duration)
    assert set(env.reset_dict) == {"time", "resolve", "noise_bef_rot", "low", "high", "radian", "rot_vel_w_pose",
                                   "vel_noise", "interp", "knee_rot"}
    t = env._termination.to_struct()
    d = HumanoidTermination(horizon=11, enable_velocity_check=True).to_struct()   # the built-in humanoid3d tables
    assert t.n_bodies == d.n_bodies == 13 and t.horizon == 11 and t.enable_velocity_check == 1
    assert t.record_all_world == 0 and t.record_world_root_pos == 0 and t.vel_offset == 136 and t.vel_divisor == 1.0
    for i in range(13):
        assert (t.body_offset[i], t.body_shape[i]) == (d.body_offset[i], d.body_shape[i])
        assert t.body_param0[i] == pytest.approx(d.body_param0[i]) and t.body_param1[i] == pytest.approx(d.body_param1[i])
    # reset: time ~ U(0, time_max) from the env's own RandomState, state from the simulator, member round-robin
    rs = np.random.RandomState(7)
    for k in range(3):
        ob = env.reset()
        time = rs.uniform(low=0, high=env.time_max)
        assert env.reset_dict["time"] == time and env.deepmimic.time == time
        clip = msg.FakeSimulator.clip
        np.testing.assert_allclose(ob, io.record_state(io.HUMANOID3D, clip.kin_pose(time), clip.kin_vel(time)), atol=0)
        assert ob.dtype == np.float64 and env.num_steps == 0 and env.reset_counter == (k + 1) % 2


def test_gym_registration_uses_the_reference_id(monkeypatch):
    """gym_simenv/__init__.py:3-6 registers 'simenv-v0'; run.py:120 makes it with deepmimic_args / dynamic_ensemble /
    reset_args keywords.  gym is not installed here: a recording stand-in for its registry shows what is registered and
    that the entry point resolves to a class accepting exactly those keywords."""
    import importlib
    import inspect
    import sys
    import types
    import amp_extensions_b200 as pkg
    seen = {}
    gym = types.ModuleType("gym")
    envs = types.ModuleType("gym.envs")
    reg = types.ModuleType("gym.envs.registration")
    reg.register = lambda id, entry_point, **kw: seen.update(id=id, entry_point=entry_point)
    gym.envs, envs.registration = envs, reg
    for name, m in (("gym", gym), ("gym.envs", envs), ("gym.envs.registration", reg)):
        monkeypatch.setitem(sys.modules, name, m)
    assert pkg.register_gym() == "simenv-v0" and seen["id"] == "simenv-v0"
    mod, cls = seen["entry_point"].split(":")
    env_cls = getattr(importlib.import_module(mod), cls)
    assert env_cls is SimEnv
    params = inspect.signature(env_cls.__init__).parameters
    for kw in ("dynamic_ensemble", "deepmimic_args", "enable_velocity_check", "horizon", "device", "seed", "reset_args"):
        assert kw in params                      # sim_env.py:27-32 (note the reference's spelling `dynamic_ensemble`)


def test_host_pipeline_chunk_bounds():
    """host_api.chunk_bounds: the chunks of a step cover [0, E) exactly once, in order, there are at most n_chunks of
    them, every cut is on a 256-row tile, and with the forward kernel's round granule (Engine.round_rows: 9 472 rows for
    4 members on 148 SMs) the cuts fall on whole rounds whenever such a split into n chunks exists."""
    from amp_extensions_b200.host_api import chunk_bounds
    assert chunk_bounds(40000, 2) == [(0, 20224), (20224, 40000)]
    assert chunk_bounds(40000, 2, 9472) == [(0, 18944), (18944, 40000)]
    assert chunk_bounds(40000, 1, 9472) == [(0, 40000)]
    assert chunk_bounds(300, 4) == [(0, 256), (256, 300)]
    assert chunk_bounds(5000, 2, 9472) == [(0, 2560), (2560, 5000)]           # shorter than a granule: plain halves
    for E in (1, 255, 256, 257, 4800, 20001, 40000, 65536, 131072, 1 << 20):
        for n in (1, 2, 3, 4, 7, 16):
            for gran in (256, 9472, 2560, 18944):     # Engine.round_rows only returns multiples of 256
                b = chunk_bounds(E, n, gran)
                assert 1 <= len(b) <= n and b[0][0] == 0 and b[-1][1] == E
                assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1)) and all(r1 > r0 for r0, r1 in b)
                assert all(r0 % 256 == 0 for r0, _ in b)
                if not all(r0 % gran == 0 for r0, _ in b):      # no aligned split with n chunks exists: the plain one
                    assert b == chunk_bounds(E, n, 256)
