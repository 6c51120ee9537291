"""GPU parity of the imitation-reward kernel against the float64 oracle (oracle/imitation_oracle.py).

Tolerance: reward and its five exponential sub-rewards within 1e-3 relative (fp32 kernel against a float64
restatement); "relative" against max(|ref|, 1e-2) because a sub-reward may underflow towards 0.
"""
import numpy as np
import pytest
import torch

from oracle import imitation_oracle as io
from tests import helpers as H

pytestmark = pytest.mark.gpu
REL = 1e-3


@pytest.fixture(scope="module", params=["h3d", "generic"])
def setup(request):
    """Both reward kernels: the register-resident humanoid3d fast path (csrc/imitation_h3d.cuh, the default for
    the built-in character) and the generic any-character kernel (csrc/imitation.cuh)."""
    import os
    from amp_extensions_b200 import ImitationReward
    clip = io.Clip(H.spinkick_raw(), io.HUMANOID3D, "wrap")
    old = os.environ.get("SIMSTEP_IMIT_GENERIC")
    os.environ["SIMSTEP_IMIT_GENERIC"] = "1" if request.param == "generic" else "0"
    try:
        imit = ImitationReward()  # the kernel is chosen when the clip is loaded
    finally:
        if old is None:
            os.environ.pop("SIMSTEP_IMIT_GENERIC", None)
        else:
            os.environ["SIMSTEP_IMIT_GENERIC"] = old
    return imit, clip


def close(x, ref, floor=1e-2, rel=REL):
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert np.isfinite(x).all()
    bad = np.abs(x - ref) > rel * np.maximum(np.abs(ref), floor)
    assert not bad.any(), f"{bad.sum()} of {bad.size} outside tolerance, worst {np.abs(x - ref).max():.3e}"


def test_clip_sampling_matches_oracle(setup):
    imit, clip = setup
    rng = np.random.default_rng(5)
    t = np.concatenate([rng.uniform(0, 5 * clip.duration, 500), clip.times[:-1], [0.0, clip.duration, 2 * clip.duration]])
    org = rng.normal(0, 0.5, (t.size, 3))
    pose, vel = imit.sample(torch.from_numpy(t).float(), torch.from_numpy(org).float())
    pose, vel = pose.cpu().numpy(), vel.cpu().numpy()
    t32 = t.astype(np.float32).astype(np.float64)
    # skip samples whose fp32 time lands on the other side of a frame / cycle boundary
    ok = np.ones(t.size, dtype=bool)
    for e in range(t.size):
        i64, _ = clip.index_blend(t[e])
        i32, _ = clip.index_blend(t32[e])
        ok[e] = i64 == i32 and clip.cycle_count(t[e]) == clip.cycle_count(t32[e])
    assert ok.sum() > 480
    ref_p = np.stack([clip.kin_pose(t32[e], org[e].astype(np.float32).astype(np.float64)) for e in range(t.size)])
    ref_v = np.stack([clip.kin_vel(t32[e]) for e in range(t.size)])
    # fp32 time resolution (6e-8 s relative) times the clip velocity bounds the achievable agreement
    np.testing.assert_allclose(pose[ok], ref_p[ok], atol=2e-5)
    np.testing.assert_allclose(vel[ok], ref_v[ok], atol=2e-3, rtol=1e-4)


def test_reward_is_one_on_the_clip_at_scale(setup):
    """1M envs (BASELINE.json config 3 size): the kinematic character scores 1 against itself."""
    imit, clip = setup
    E = 1_000_000
    g = torch.Generator(device="cuda").manual_seed(7)
    t = torch.rand(E, device="cuda", generator=g) * (6 * clip.duration)
    pose, vel = imit.sample(t)
    r, terms = imit.reward(pose, vel, t, want_terms=True)
    assert float((r - 1).abs().max()) < 2e-5
    assert float((terms - 1).abs().max()) < 1e-4
    # rewards are bounded and fall when the pose is perturbed
    r2 = imit.reward(pose + 0.05 * torch.randn(pose.shape, device="cuda", generator=g), vel, t)
    assert float(r2.max()) <= 1.0 + 1e-6 and float(r2.min()) >= 0.0 and float(r2.mean()) < float(r.mean())


@pytest.mark.parametrize("with_origin", [False, True])
def test_reward_matches_oracle_on_perturbed_poses(setup, with_origin):
    imit, clip = setup
    E = 384
    pose, vel, t, origin = H.perturbed_poses(E, seed=3, clip=clip, t_max=4 * clip.duration, with_origin=with_origin)
    t32 = t.astype(np.float32).astype(np.float64)
    p32, v32 = pose.astype(np.float32).astype(np.float64), vel.astype(np.float32).astype(np.float64)
    o32 = origin.astype(np.float32).astype(np.float64) if with_origin else None
    ref_r, ref_terms = io.imitation_reward_batch(io.HUMANOID3D, clip, p32, v32, t32, o32)
    r, terms = imit.reward(torch.from_numpy(pose).float(), torch.from_numpy(vel).float(), torch.from_numpy(t).float(),
                           torch.from_numpy(origin).float() if with_origin else None, want_terms=True)
    close(terms.cpu().numpy(), ref_terms)
    close(r.cpu().numpy(), ref_r)
    assert ref_terms.min() < 0.5 < ref_terms.max()  # the inputs exercise the exponentials' range


def test_closed_form_terms_on_gpu(setup):
    imit, clip = setup
    t = 0.45
    p1, v1 = clip.kin_pose(t), clip.kin_vel(t)
    theta = 0.3
    ax = np.array([0.2, 1.0, -0.4])
    dq = np.concatenate([[np.cos(theta / 2)], np.sin(theta / 2) * ax / np.linalg.norm(ax)])
    rows_p, rows_v, want = [], [], []
    p0 = p1.copy(); p0[7:11] = io.quat_mul(p1[7:11], dq)
    rows_p.append(p0); rows_v.append(v1); want.append((0, np.exp(-2.0 * (0.5 / 4.8) * theta ** 2)))
    p0 = p1.copy(); p0[19] += 0.2
    rows_p.append(p0); rows_v.append(v1); want.append((0, np.exp(-2.0 * (0.3 / 4.8) * 0.04)))
    v0 = v1.copy(); v0[11:14] += [0.5, -0.25, 1.0]
    rows_p.append(p1); rows_v.append(v0); want.append((1, np.exp(-0.1 * (0.3 / 4.8) * 1.3125)))
    p0 = p1.copy(); p0[0] += 0.1; p0[2] -= 0.2
    rows_p.append(p0); rows_v.append(v1); want.append((3, np.exp(-5.0 * 0.05)))
    _, terms = imit.reward(torch.tensor(np.stack(rows_p)).float(), torch.tensor(np.stack(rows_v)).float(),
                           torch.full((4,), t), want_terms=True)
    terms = terms.cpu().numpy()
    for row, (k, val) in enumerate(want):
        assert terms[row, k] == pytest.approx(val, rel=REL)


def test_unaligned_rows_and_ragged_tiles(setup):
    """Row blocks that are not 16-byte aligned (no TMA bulk copy) and batch sizes that leave ragged tiles."""
    imit, clip = setup
    pose, vel, t, origin = H.perturbed_poses(301, seed=11, clip=clip)
    args = [torch.from_numpy(a).float().cuda() for a in (pose, vel, t, origin)]
    r = imit.reward(*args)
    for lo, hi in ((1, 301), (3, 132), (0, 129), (2, 3), (0, 257)):
        sub = [a[lo:hi].clone() for a in args]                        # aligned copies
        assert torch.equal(imit.reward(*sub), r[lo:hi])
    # misaligned base pointers: views into a buffer shifted by one float
    buf_p = torch.empty(pose.size + 1, device="cuda"); buf_v = torch.empty(vel.size + 1, device="cuda")
    vp = buf_p[1:].view(301, 43); vv = buf_v[1:].view(301, 43)
    vp.copy_(args[0]); vv.copy_(args[1])
    eng = imit.engine
    import ctypes as C
    out = torch.empty(301, device="cuda")
    rc = eng.lib.simstep_imitation_reward(eng._h, C.c_void_p(vp.data_ptr()), C.c_void_p(vv.data_ptr()),
                                          C.c_void_p(args[2].data_ptr()), C.c_void_p(args[3].data_ptr()), 301,
                                          C.c_void_p(out.data_ptr()), None, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(out, r)


def test_rows_are_independent(setup):
    imit, clip = setup
    pose, vel, t, origin = H.perturbed_poses(300, seed=9, clip=clip)
    args = [torch.from_numpy(a).float().cuda() for a in (pose, vel, t, origin)]
    r = imit.reward(*args)
    perm = torch.randperm(300, device="cuda")
    assert torch.equal(imit.reward(*[a[perm] for a in args]), r[perm])
    assert torch.equal(imit.reward(*[a[100:133] for a in args]), r[100:133])
    assert imit.reward(*[a[:0] for a in args]).shape == (0,)


# ---------------------------------------------------------------------------------------------------
# record_state: the env's state vector from generalized pose / velocity (CtController.cpp:378-495)


@pytest.mark.parametrize("flags", [dict(), dict(record_world_root_rot=False), dict(record_all_world=True),
                                   dict(record_world_root_pos=True, vel_scale=1.0 / 30.0)])
def test_record_state_matches_oracle(setup, flags):
    """fp32 kernel against the float64 literal restatement (4x4 transforms + spatial Jacobian velocities); scale of
    a position / direction feature is 1, of a velocity feature the sample's own spread."""
    imit, clip = setup
    pose, vel, _, _ = H.perturbed_poses(64, seed=7, clip=clip, with_origin=False)
    st = imit.record_state(torch.from_numpy(pose).float(), torch.from_numpy(vel).float(), **flags).cpu().numpy()
    ref = np.stack([io.record_state(io.HUMANOID3D, pose[e], vel[e], **flags) for e in range(64)])
    assert st.shape == (64, 226)
    np.testing.assert_allclose(st[:, :136], ref[:, :136], rtol=0, atol=REL * 1.0 * 0.05)       # 5e-5 absolute
    vscale = float(np.abs(ref[:, 136:]).mean())
    np.testing.assert_allclose(st[:, 136:], ref[:, 136:], rtol=0, atol=REL * max(vscale, 1e-3))


def test_clip_reset_states_feed_the_env(setup):
    """VecSimEnv.clip_reset_fn: envs start on the reference motion; the states are what record_state gives for the
    sampled clip poses, upright (no fall contact) and within the clip's height range."""
    from amp_extensions_b200 import AmpDataset, DynamicsEnsemble, VecSimEnv
    imit, clip = setup
    t = torch.linspace(0.0, float(clip.duration) * 0.999, 50)
    states = imit.reset_states(t.cuda()).cpu().numpy()
    ref = np.stack([io.record_state(io.HUMANOID3D, clip.kin_pose(float(x)), clip.kin_vel(float(x))) for x in t])
    np.testing.assert_allclose(states[:, :136], ref[:, :136], rtol=0, atol=2e-4)
    np.testing.assert_allclose(states[:, 136:], ref[:, 136:], rtol=0, atol=2e-3 * max(1.0, np.abs(ref[:, 136:]).max()))
    s, a, s2 = H.synth_dataset(512, 226, 28, 0)
    ens = DynamicsEnsemble(226, 28, AmpDataset(s, a, s2), None, num_models=2, hidden_sizes=[64, 64], dense_connect=True,
                           transform=True, base_seed=100)
    env = VecSimEnv(ens, 32, reset_fn=VecSimEnv.clip_reset_fn(imit), seed=3)
    ob = env.reset().cpu().numpy()
    assert ob.shape == (32, 226) and np.isfinite(ob).all()
    assert (ob[:, 0] > 0.6).all() and (ob[:, 0] < 1.3).all()           # root height along the spin kick
    from oracle import milo_oracle as mo
    assert not mo.simenv_collided(ob.astype(np.float64)).any()          # the reference motion never falls


def test_reset_noise_states_match_the_oracle(setup):
    """reset_states(reset_args=...) = clip sample -> cKinCharacter::AddNoise (KinCharacter.cpp:340-532, batched on the
    device in reset_noise.add_reset_noise) -> record_state, against the float64 oracle fed the same draws (the oracle
    itself is pinned to the reference's compiled primitives in tests/test_imitation_ref.py); then the plugin's own call:
    VecSimEnv.clip_reset_fn(reset_args=<the reference's dict>) gives finite, upright, perturbed states and honours
    custom_time."""
    from amp_extensions_b200 import AmpDataset, DynamicsEnsemble, VecSimEnv
    from amp_extensions_b200 import reset_noise as rn
    imit, clip = setup
    rng = np.random.default_rng(11)
    E, dof = 40, 43
    ra = dict(custom_time=False, time_min=0, time_max=0, resolve=True, noise_bef_rot=True, noise_min=-0.01,
              noise_max=0.02, radian=0.25, rot_vel_w_pose=True, vel_noise=True, interp=0.6, knee_rot=False)
    kw = rn.reset_kwargs(ra)
    n_r = rn.num_rotation_draws(imit.character, kw["vel_noise"], kw["knee_rot"])
    t = np.linspace(0.0, float(clip.duration) * 0.999, E).astype(np.float32)
    u_pose = rng.uniform(0, 1, (E, dof)).astype(np.float32)
    u_vel = rng.uniform(0, 1, (E, dof)).astype(np.float32)
    r = rng.uniform(-1, 1, (E, n_r)).astype(np.float32)
    states = imit.reset_states(torch.from_numpy(t).cuda(), reset_args=ra, draws=(u_pose, u_vel, r)).cpu().numpy()
    ref = []
    for e in range(E):
        p, v, used = io.reset_noise(io.HUMANOID3D, clip.kin_pose(float(t[e])), clip.kin_vel(float(t[e])),
                                    u_pose[e].astype(np.float64), u_vel[e].astype(np.float64), r[e].astype(np.float64), **kw)
        assert used == n_r
        ref.append(io.record_state(io.HUMANOID3D, p, v))
    ref = np.stack(ref)
    np.testing.assert_allclose(states[:, :136], ref[:, :136], rtol=0, atol=3e-4)
    np.testing.assert_allclose(states[:, 136:], ref[:, 136:], rtol=0, atol=2e-3 * max(1.0, np.abs(ref[:, 136:]).max()))
    plain = imit.reset_states(torch.from_numpy(t).cuda()).cpu().numpy()
    assert np.abs(states - plain).max() > 1e-2                        # the noise did something

    s, a, s2 = H.synth_dataset(512, 226, 28, 0)
    ens = DynamicsEnsemble(226, 28, AmpDataset(s, a, s2), None, num_models=2, hidden_sizes=[64, 64], dense_connect=True,
                           transform=True, base_seed=100)
    ra2 = dict(ra, custom_time=True, time_min=0.2, time_max=0.2, noise_min=0.0, noise_max=0.0, radian=0.05)
    env = VecSimEnv(ens, 32, reset_fn=VecSimEnv.clip_reset_fn(imit, reset_args=ra2), seed=3)
    ob = env.reset().cpu().numpy()
    assert ob.shape == (32, 226) and np.isfinite(ob).all()
    assert (ob[:, 0] > 0.6).all() and (ob[:, 0] < 1.3).all()
    at = imit.reset_states(torch.full((1,), 0.2).cuda()).cpu().numpy()[0]
    d = np.abs(ob[:, :136] - at[:136]).max()                           # all at t = 0.2 s, each perturbed a little (0.05 rad);
    assert 1e-4 < d < 0.3                                               # the velocities are also scaled by interp = 0.6
    assert np.abs(ob[0] - ob[1]).max() > 1e-4
    ob2 = VecSimEnv(ens, 32, reset_fn=VecSimEnv.clip_reset_fn(imit, reset_args=ra2), seed=3).reset().cpu().numpy()
    np.testing.assert_array_equal(ob, ob2)                             # seeded: reproducible


# ---- against outputs of the reference's own compiled kinematics code (tests/golden/imitation_ref_golden.npz) ----

def _ref_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "imitation_ref_golden.npz"))


@pytest.mark.parametrize("tag", ["plain", "origin", "far"])
def test_reward_matches_reference_golden(setup, tag):
    """The CUDA reward against cKinTree / cRBDUtil / cMotion themselves (compiled from the reference sources by
    oracle/ref_build.py, vectors written by tests/golden/make_imitation_ref_golden.py).  Inputs are exactly
    float32-representable; tolerance 1e-3 relative on the reward and on each exponential sub-reward."""
    imit, _ = setup
    g = _ref_golden()
    pose, vel, t = g[f"{tag}_pose"], g[f"{tag}_vel"], g[f"{tag}_t"]
    origin = torch.from_numpy(g[f"{tag}_origin"]).float() if tag == "origin" else None
    r, terms = imit.reward(torch.from_numpy(pose).float(), torch.from_numpy(vel).float(), torch.from_numpy(t).float(),
                           origin, want_terms=True)
    close(terms.cpu().numpy(), g[f"{tag}_terms"])
    close(r.cpu().numpy(), g[f"{tag}_reward"])


def test_clip_sampling_matches_reference_golden(setup):
    imit, clip = setup
    g = _ref_golden()
    t, org = g["sample_t"], g["sample_origin"]
    pose, vel = imit.sample(torch.from_numpy(t).float(), torch.from_numpy(org).float())
    # a sample whose time sits on a cycle boundary may wrap one way in fp32 and the other in float64
    phase = np.mod(t, clip.duration)
    ok = np.minimum(phase, clip.duration - phase) > 1e-4
    assert ok.sum() >= t.size - 4
    np.testing.assert_allclose(pose.cpu().numpy()[ok], g["sample_pose"][ok], atol=2e-5)
    np.testing.assert_allclose(vel.cpu().numpy()[ok], g["sample_vel"][ok], atol=2e-3, rtol=1e-4)


def test_record_state_matches_reference_golden(setup):
    imit, _ = setup
    g = _ref_golden()
    pose, vel = g["plain_pose"], g["plain_vel"]
    n = g["state_features"].shape[1]
    for k, (aw, wrp, wrr, vs) in enumerate(g["state_flags"]):
        st = imit.record_state(torch.from_numpy(pose[:n]).float(), torch.from_numpy(vel[:n]).float(),
                               record_all_world=bool(aw), record_world_root_pos=bool(wrp),
                               record_world_root_rot=bool(wrr), vel_scale=float(vs)).cpu().numpy()
        ref = g["state_features"][k]
        np.testing.assert_allclose(st[:, :136], ref[:, :136], rtol=0, atol=5e-5)
        vscale = float(np.abs(ref[:, 136:]).mean())
        np.testing.assert_allclose(st[:, 136:], ref[:, 136:], rtol=0, atol=REL * max(vscale, 1e-3))


@pytest.mark.parametrize("name", ["humanoid3d_kick", "humanoid3d_run"])
def test_other_clips_match_reference_golden(name):
    """A non-looping reference clip (kick: clamped index, zero velocity past the end, no cycle offset) and a short
    cycle (run), loaded by the product's own MotionClip.from_raw and evaluated by both reward kernels, against the
    compiled reference (tests/golden/imitation_ref_golden.npz)."""
    import os
    from amp_extensions_b200 import ImitationReward
    from amp_extensions_b200.character import humanoid3d
    from amp_extensions_b200.motion import MotionClip
    g = _ref_golden()
    raw, loop = g[f"{name}/raw"], str(g[f"{name}/loop"])
    dur = float(g[f"{name}/duration"])
    for generic in ("0", "1"):
        old = os.environ.get("SIMSTEP_IMIT_GENERIC")
        os.environ["SIMSTEP_IMIT_GENERIC"] = generic
        try:
            ch = humanoid3d()
            imit = ImitationReward(character=ch, clip=MotionClip.from_raw(raw, ch, loop))
        finally:
            if old is None:
                os.environ.pop("SIMSTEP_IMIT_GENERIC", None)
            else:
                os.environ["SIMSTEP_IMIT_GENERIC"] = old
        t, org = g[f"{name}/sample_t"], g[f"{name}/sample_origin"]
        pose, vel = imit.sample(torch.from_numpy(t).float(), torch.from_numpy(org).float())
        # away from frame-0 / end-of-clip boundaries, where fp32 time may fall on the other side
        phase = np.mod(t, dur)
        ok = (np.minimum(phase, dur - phase) > 1e-4) & (np.abs(t) > 1e-4)
        assert ok.sum() >= t.size - 4
        np.testing.assert_allclose(pose.cpu().numpy()[ok], g[f"{name}/sample_pose"][ok], atol=2e-5)
        np.testing.assert_allclose(vel.cpu().numpy()[ok], g[f"{name}/sample_vel"][ok], atol=2e-3, rtol=1e-4)
        r, terms = imit.reward(torch.from_numpy(g[f"{name}/pose"]).float(), torch.from_numpy(g[f"{name}/vel"]).float(),
                               torch.from_numpy(g[f"{name}/t"]).float(), None, want_terms=True)
        close(terms.cpu().numpy(), g[f"{name}/terms"])
        close(r.cpu().numpy(), g[f"{name}/reward"])


def test_a_different_skeleton_matches_reference_golden():
    """The any-character kernel on the reference's dog3d (23 joints, 83 dof, four end effectors) and its trot clip,
    loaded by the product's own Character.from_json / MotionClip.from_raw, against the compiled reference."""
    import json
    from amp_extensions_b200 import ImitationReward
    from amp_extensions_b200.character import Character
    from amp_extensions_b200.motion import MotionClip
    g = _ref_golden()
    ch = Character.from_json(json.loads(str(g["dog/character_json"])), ignore_body_rotation=True)
    imit = ImitationReward(character=ch, clip=MotionClip.from_raw(g["dog/raw"], ch, str(g["dog/loop"])))
    dur = float(g["dog/duration"])
    t, org = g["dog/sample_t"], g["dog/sample_origin"]
    pose, vel = imit.sample(torch.from_numpy(t).float(), torch.from_numpy(org).float())
    phase = np.mod(t, dur)
    ok = np.minimum(phase, dur - phase) > 1e-4
    np.testing.assert_allclose(pose.cpu().numpy()[ok], g["dog/sample_pose"][ok], atol=2e-5)
    np.testing.assert_allclose(vel.cpu().numpy()[ok], g["dog/sample_vel"][ok], atol=2e-3, rtol=1e-4)
    r, terms = imit.reward(torch.from_numpy(g["dog/pose"]).float(), torch.from_numpy(g["dog/vel"]).float(),
                           torch.from_numpy(g["dog/t"]).float(), None, want_terms=True)
    close(terms.cpu().numpy(), g["dog/terms"])
    close(r.cpu().numpy(), g["dog/reward"])
    with pytest.raises(NotImplementedError):
        imit.record_state(torch.from_numpy(g["dog/pose"]).float(), torch.from_numpy(g["dog/vel"]).float())
