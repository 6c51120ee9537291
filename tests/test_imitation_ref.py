"""CPU: the float64 restatement (oracle/imitation_oracle.py) pinned to the reference's OWN DeepMimicCore code.

Two layers:
  * tests/golden/imitation_ref_golden.npz holds outputs of the reference's MathUtil / KinTree / Motion / RBDUtil /
    SpAlg sources compiled where they lie (oracle/ref_build.py; Eigen 3.3.7 replaced by oracle/eigen_shim) — these
    comparisons run everywhere, including the GPU box, which has no reference tree;
  * when oracle/_ref/libdmref.so can be built or was shipped, the same functions are also compared live on fresh
    random inputs.
Tolerance: both sides are float64 and differ only in summation order: 1e-9 absolute on O(1) quantities.
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import imitation_oracle as io
from oracle import ref_build as rb
from tests import helpers as H

CH = io.HUMANOID3D
ATOL = 1e-9
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "imitation_ref_golden.npz"))
PD = ctypes.POINTER(ctypes.c_double)


def P(a):
    return a.ctypes.data_as(PD)


@pytest.fixture(scope="module")
def clip():
    return io.Clip(H.spinkick_raw(), CH, "wrap")


@pytest.fixture(scope="module")
def lib():
    handle = rb.load()
    if handle is None:
        pytest.skip("oracle/_ref/libdmref.so is not available on this machine")
    return handle


# ---- against the committed vectors ---------------------------------------------------------------------

def test_character_tables_match_the_reference_loader():
    offs, sizes = io.param_layout(CH)
    assert list(GOLD["param_offset"]) == offs and list(GOLD["param_size"]) == sizes
    w = np.asarray(CH["diff_weight"], dtype=np.float64)
    np.testing.assert_allclose(GOLD["joint_weights"], w / np.abs(w).sum(), atol=1e-15)


def test_clip_tables_match_the_reference_loader(clip):
    """cMotion::Load + PostProcessFrames + BuildFrameVel, cKinController::PostProcessMotion."""
    assert int(GOLD["clip_loop"]) == 1 and clip.loop
    assert float(GOLD["clip_duration"]) == pytest.approx(clip.duration, abs=1e-15)
    np.testing.assert_allclose(clip.times, GOLD["clip_times"], atol=1e-15)
    np.testing.assert_allclose(clip.frames, GOLD["clip_frames"], atol=1e-14)
    np.testing.assert_allclose(clip.vels, GOLD["clip_vels"], atol=1e-11)


def test_clip_sampling_matches_the_reference(clip):
    """cMotion::CalcFrame / CalcFrameVel (slerp blend, wrap, negative times) + loop root offset + origin."""
    t, org = GOLD["sample_t"], GOLD["sample_origin"]
    for e in range(t.size):
        np.testing.assert_allclose(clip.kin_pose(float(t[e]), org[e]), GOLD["sample_pose"][e], atol=ATOL)
        np.testing.assert_allclose(clip.kin_vel(float(t[e])), GOLD["sample_vel"][e], atol=ATOL)


@pytest.mark.parametrize("tag", ["plain", "origin", "far"])
def test_reward_matches_the_reference(clip, tag):
    pose, vel, t = GOLD[f"{tag}_pose"], GOLD[f"{tag}_vel"], GOLD[f"{tag}_t"]
    origin = GOLD[f"{tag}_origin"] if tag == "origin" else None
    n = 48
    r, terms = io.imitation_reward_batch(CH, clip, pose[:n], vel[:n], t[:n], None if origin is None else origin[:n])
    np.testing.assert_allclose(terms, GOLD[f"{tag}_terms"][:n], atol=ATOL)
    np.testing.assert_allclose(r, GOLD[f"{tag}_reward"][:n], atol=ATOL)


def test_building_blocks_match_the_reference(clip):
    """CalcPoseErr / CalcVelErr per joint, forward kinematics, heading, origin transform, centre of mass."""
    pose, vel, t = GOLD["plain_pose"], GOLD["plain_vel"], GOLD["plain_t"]
    E, nj = GOLD["blk_pose_err"].shape
    for e in range(E):
        p1, v1 = clip.kin_pose(float(t[e])), clip.kin_vel(float(t[e]))
        th = io.quat_theta(io.quat_diff(pose[e, 3:7], p1[3:7]))
        assert th * th == pytest.approx(GOLD["blk_pose_err"][e, 0], abs=ATOL)
        d = v1[3:7] - vel[e, 3:7]
        assert float(d @ d) == pytest.approx(GOLD["blk_vel_err"][e, 0], abs=ATOL)
        for j in range(1, nj):
            assert io.calc_pose_err(CH, j, pose[e], p1) == pytest.approx(GOLD["blk_pose_err"][e, j], abs=ATOL)
            assert io.calc_vel_err(CH, j, vel[e], v1) == pytest.approx(GOLD["blk_vel_err"][e, j], abs=ATOL)
        for j in range(nj):
            np.testing.assert_allclose(io.calc_joint_world_pos(CH, pose[e], j), GOLD["blk_joint_pos"][e, j], atol=ATOL)
            np.testing.assert_allclose(io.joint_world_trans(CH, pose[e], j).reshape(16), GOLD["blk_joint_trans"][e, j],
                                       atol=ATOL)
        assert io.calc_heading(pose[e, 3:7]) == pytest.approx(GOLD["blk_heading"][e], abs=ATOL)
        np.testing.assert_allclose(io.build_origin_trans(pose[e]).reshape(16), GOLD["blk_origin_trans"][e], atol=ATOL)
        com, com_vel = io.calc_com(CH, pose[e], vel[e])
        np.testing.assert_allclose(com, GOLD["blk_com"][e], atol=ATOL)
        np.testing.assert_allclose(com_vel, GOLD["blk_com_vel"][e], atol=ATOL)


def test_state_features_match_the_reference():
    """cCtController::BuildStatePose / BuildStateVel laid over the reference's BodyWorldTrans / RotMatToQuaternion /
    CalcNormalTangent / CalcBodyPartVel / CalcJointWorldAngularVel — a different route from the restatement's
    rotation matrices and spatial Jacobian."""
    pose, vel = GOLD["plain_pose"], GOLD["plain_vel"]
    for k, (aw, wrp, wrr, vs) in enumerate(GOLD["state_flags"]):
        for e in range(GOLD["state_features"].shape[1]):
            st = io.record_state(CH, pose[e], vel[e], record_all_world=bool(aw), record_world_root_pos=bool(wrp),
                                 record_world_root_rot=bool(wrr), vel_scale=float(vs))
            np.testing.assert_allclose(st, GOLD["state_features"][k, e], atol=ATOL)


# ---- live against the compiled reference ------------------------------------------------------------------

def test_reference_library_exports(lib):
    for name in ("dmref_init", "dmref_clip_table", "dmref_kin_pose_vel", "dmref_pose_err", "dmref_vel_err",
                 "dmref_joint_world_pos", "dmref_joint_world_trans", "dmref_heading", "dmref_origin_trans", "dmref_com",
                 "dmref_lerp_poses", "dmref_calc_vel", "dmref_quat_theta", "dmref_quat_rot_vec", "dmref_normal_tangent",
                 "dmref_reward", "dmref_reward_batch", "dmref_record_state"):
        assert hasattr(lib, name)
    assert lib.dmref_num_dof() == 43 and lib.dmref_num_joints() == 15 and lib.dmref_num_frames() == 78


def test_live_reward_on_fresh_inputs(lib, clip):
    rng = np.random.default_rng(2024)
    E = 40
    pose, vel, t, origin = H.perturbed_poses(E, seed=99, clip=clip, t_max=5 * clip.duration, with_origin=True)
    t[:6] = [-0.7, -1e-9, 0.0, clip.duration, 3 * clip.duration, clip.times[5]]   # boundaries of the wrap / blend logic
    pose[::3, 0:3] += rng.normal(0, 0.3, (len(pose[::3]), 3))
    pose, vel, t, origin = (np.ascontiguousarray(x) for x in (pose, vel, t, origin))
    r, terms = np.zeros(E), np.zeros((E, 5))
    lib.dmref_reward_batch(E, P(pose), P(vel), P(t), P(origin), P(r), P(terms))
    ro, to = io.imitation_reward_batch(CH, clip, pose, vel, t, origin)
    np.testing.assert_allclose(to, terms, atol=ATOL)
    np.testing.assert_allclose(ro, r, atol=ATOL)


def test_live_quaternion_helpers(lib):
    """QuatTheta's branches (w > 1 renormalised, sin < 1e-4 -> 0), q*v, slerp through LerpPoses incl. the
    |dot| >= 1 - eps linear branch and the negative-dot flip, CalcVel's axis-angle velocities."""
    rng = np.random.default_rng(5)
    for k in range(200):
        q = rng.normal(0, 1, 4)
        q /= np.linalg.norm(q)
        if k % 10 == 0:
            q = np.array([1.0 + 1e-9 * k, 1e-6, 0, 0])
        if k % 10 == 1:
            q = np.array([np.cos(2e-5), np.sin(2e-5), 0, 0])
        assert io.quat_theta(q) == pytest.approx(lib.dmref_quat_theta(P(np.ascontiguousarray(q))), abs=1e-12)
        v, out = rng.normal(0, 1, 3), np.zeros(3)
        lib.dmref_quat_rot_vec(P(np.ascontiguousarray(q)), P(v), P(out))
        np.testing.assert_allclose(io.quat_rot_vec(q, v), out, atol=1e-12)
    clip = io.Clip(H.spinkick_raw(), CH, "wrap")
    offs, sizes = io.param_layout(CH)
    for k in range(40):
        a, b = clip.frames[rng.integers(0, 78)].copy(), clip.frames[rng.integers(0, 78)].copy()
        if k % 4 == 0:
            b = a.copy()                      # identical quaternions: the linear branch of slerp
        if k % 4 == 1:
            b[offs[1]:offs[1] + 4] *= -1      # same rotation, opposite sign: the flip
        lerp = float(rng.uniform(0, 1))
        out = np.zeros(43)
        lib.dmref_lerp_poses(P(a), P(b), lerp, P(out))
        np.testing.assert_allclose(io.lerp_poses(CH, a, b, lerp), out, atol=1e-12)
        lib.dmref_calc_vel(P(a), P(b), 1.0 / 60, P(out))
        np.testing.assert_allclose(io.calc_vel(CH, a, b, 1.0 / 60), out, atol=1e-9)


# ---- other clips of the reference: a non-looping motion and a short cycle -----------------------------------------

EXTRA_CLIPS = ["humanoid3d_kick", "humanoid3d_run"]


@pytest.mark.parametrize("name", EXTRA_CLIPS)
def test_other_clips_match_the_reference(name):
    """cMotion::Load / CalcFrame / CalcFrameVel with Loop "none" (index clamp, no cycle offset) and "wrap", and the
    reward against those clips; the restatement AND the product's host-side loader (motion.MotionClip.from_raw, the
    tables it uploads) against what the compiled reference holds after loading the same file."""
    from amp_extensions_b200.character import humanoid3d
    from amp_extensions_b200.motion import MotionClip
    raw, loop = GOLD[f"{name}/raw"], str(GOLD[f"{name}/loop"])
    assert loop == ("none" if name == "humanoid3d_kick" else "wrap")
    c = io.Clip(raw, CH, loop)
    mc = MotionClip.from_raw(raw, humanoid3d(), loop)
    for frames, vels, times, dur in ((c.frames, c.vels, c.times, c.duration),
                                     (mc.frames, mc.frame_vels, mc.frame_times, mc.duration)):
        np.testing.assert_allclose(times, GOLD[f"{name}/times"], atol=1e-14)
        np.testing.assert_allclose(frames, GOLD[f"{name}/frames"], atol=1e-13)
        np.testing.assert_allclose(vels, GOLD[f"{name}/vels"], atol=1e-10)
        assert dur == pytest.approx(float(GOLD[f"{name}/duration"]), abs=1e-14)
    assert mc.loop_wrap == (loop == "wrap")
    np.testing.assert_allclose(mc.cycle_delta, c.cycle_delta, atol=1e-14)
    t, org = GOLD[f"{name}/sample_t"], GOLD[f"{name}/sample_origin"]
    for e in range(t.size):
        np.testing.assert_allclose(c.kin_pose(float(t[e]), org[e]), GOLD[f"{name}/sample_pose"][e], atol=ATOL)
        np.testing.assert_allclose(c.kin_vel(float(t[e])), GOLD[f"{name}/sample_vel"][e], atol=ATOL)
    n = 24
    r, terms = io.imitation_reward_batch(CH, c, GOLD[f"{name}/pose"][:n], GOLD[f"{name}/vel"][:n], GOLD[f"{name}/t"][:n])
    np.testing.assert_allclose(terms, GOLD[f"{name}/terms"][:n], atol=ATOL)
    np.testing.assert_allclose(r, GOLD[f"{name}/reward"][:n], atol=ATOL)


def test_a_different_skeleton_matches_the_reference():
    """The reference's dog3d character (23 joints, 83 dof, four end effectors, tails) and its trot clip through the
    same restatement and through the product's Character.from_json / MotionClip.from_raw host loaders, against the
    compiled reference initialised with the same two files."""
    import json
    from amp_extensions_b200.character import Character
    from amp_extensions_b200.motion import MotionClip
    cj = json.loads(str(GOLD["dog/character_json"]))
    dog = H.character_dict_from_json(cj)
    raw, loop = GOLD["dog/raw"], str(GOLD["dog/loop"])
    c = io.Clip(raw, dog, loop)
    with pytest.raises(NotImplementedError):
        Character.from_json(cj)                      # the neck and tail bodies carry attach rotations
    ch = Character.from_json(cj, ignore_body_rotation=True)
    assert ch.dof == 83 and ch.n_joints == 23 and ch.body_rotation_ignored
    np.testing.assert_allclose(ch.joint_weights(), GOLD["dog/joint_weights"], atol=1e-15)
    mc = MotionClip.from_raw(raw, ch, loop)
    for frames, vels, times in ((c.frames, c.vels, c.times), (mc.frames, mc.frame_vels, mc.frame_times)):
        np.testing.assert_allclose(times, GOLD["dog/times"], atol=1e-14)
        np.testing.assert_allclose(frames, GOLD["dog/frames"], atol=1e-13)
        np.testing.assert_allclose(vels, GOLD["dog/vels"], atol=1e-10)
    t, org = GOLD["dog/sample_t"], GOLD["dog/sample_origin"]
    for e in range(t.size):
        np.testing.assert_allclose(c.kin_pose(float(t[e]), org[e]), GOLD["dog/sample_pose"][e], atol=ATOL)
        np.testing.assert_allclose(c.kin_vel(float(t[e])), GOLD["dog/sample_vel"][e], atol=ATOL)
    for e in range(GOLD["dog/com"].shape[0]):
        com, com_vel = io.calc_com(dog, GOLD["dog/pose"][e], GOLD["dog/vel"][e])
        np.testing.assert_allclose(com, GOLD["dog/com"][e], atol=ATOL)
        np.testing.assert_allclose(com_vel, GOLD["dog/com_vel"][e], atol=ATOL)
    n = 12
    r, terms = io.imitation_reward_batch(dog, c, GOLD["dog/pose"][:n], GOLD["dog/vel"][:n], GOLD["dog/t"][:n])
    np.testing.assert_allclose(terms, GOLD["dog/terms"][:n], atol=ATOL)
    np.testing.assert_allclose(r, GOLD["dog/reward"][:n], atol=ATOL)


# ---- reset noise (KinCharacter.cpp:340-532) ---------------------------------------------------------------

RESET_CASES = [
    dict(noise_bef_rot=False, noise_min=0.0, noise_max=0.0, radian=0.3, rot_vel_w_pose=False, vel_noise=False, interp=1.0, knee_rot=False),
    dict(noise_bef_rot=True, noise_min=-0.05, noise_max=0.1, radian=0.5, rot_vel_w_pose=True, vel_noise=True, interp=0.4, knee_rot=True),
    dict(noise_bef_rot=False, noise_min=-0.02, noise_max=0.02, radian=1.2, rot_vel_w_pose=False, vel_noise=True, interp=0.0, knee_rot=False),
    dict(noise_bef_rot=True, noise_min=0.01, noise_max=0.03, radian=0.0, rot_vel_w_pose=True, vel_noise=True, interp=0.5, knee_rot=True),
    dict(noise_bef_rot=False, noise_min=0.0, noise_max=0.0, radian=1e-7, rot_vel_w_pose=True, vel_noise=False, interp=1.0, knee_rot=True),
]


@pytest.mark.parametrize("case", range(len(RESET_CASES)))
def test_reset_noise_matches_the_reference_primitives(lib, clip, case):
    """oracle.reset_noise against dmref_reset_noise: cKinCharacter::AddNoise's loop restated in the driver around the
    reference's own compiled cMathUtil / cKinTree calls (KinCharacter.cpp itself needs OpenGL headers), fed the same
    draws: every flag combination incl. radian == 0 (no rotation, no interpolation), a vanishing rotation (the
    |theta| < 1e-5 branch of EulerToAxisAngle), the knee rules and the `!(j == 4 || j != 10)` condition of the
    velocity noise."""
    kw = RESET_CASES[case]
    rng = np.random.default_rng(100 + case)
    dof = 43
    for trial in range(6):
        t = float(rng.uniform(0, clip.duration))
        pose, vel = clip.kin_pose(t), clip.kin_vel(t)
        u_pose, u_vel = rng.uniform(0, 1, dof), rng.uniform(0, 1, dof)
        r = rng.uniform(-1, 1, 64)
        want_p, want_v = np.zeros(dof), np.zeros(dof)
        used = lib.dmref_reset_noise(P(np.ascontiguousarray(pose)), P(np.ascontiguousarray(vel)), int(kw["noise_bef_rot"]),
                                     kw["noise_min"], kw["noise_max"], kw["radian"], int(kw["rot_vel_w_pose"]),
                                     int(kw["vel_noise"]), kw["interp"], int(kw["knee_rot"]), P(u_pose), P(u_vel), P(r),
                                     P(want_p), P(want_v))
        got_p, got_v, got_used = io.reset_noise(CH, pose, vel, u_pose, u_vel, r, **kw)
        assert got_used == used
        np.testing.assert_allclose(got_p, want_p, atol=ATOL)
        np.testing.assert_allclose(got_v, want_v, atol=ATOL)
        if kw["radian"] == 0 and kw["noise_min"] == 0 and kw["noise_max"] == 0:
            np.testing.assert_array_equal(got_p, pose)


def test_reset_noise_product_matches_the_oracle_on_cpu():
    """The batched torch implementation the product uses (amp_extensions_b200/reset_noise.py; plain tensor ops, so it
    also runs on the CPU) against the float64 restatement with the same draws, all flag sets, float64 tensors."""
    import torch
    from amp_extensions_b200 import reset_noise as rn
    from amp_extensions_b200.character import humanoid3d
    ch = humanoid3d()
    c = io.Clip(H.spinkick_raw(), CH, "wrap")
    rng = np.random.default_rng(7)
    E, dof = 9, 43
    for kw in RESET_CASES:
        n_r = rn.num_rotation_draws(ch, kw["vel_noise"], kw["knee_rot"])
        ts = rng.uniform(0, c.duration, E)
        pose = np.stack([c.kin_pose(float(t)) for t in ts])
        vel = np.stack([c.kin_vel(float(t)) for t in ts])
        u_pose, u_vel, r = rng.uniform(0, 1, (E, dof)), rng.uniform(0, 1, (E, dof)), rng.uniform(-1, 1, (E, n_r))
        got_p, got_v = rn.add_reset_noise(ch, torch.from_numpy(pose), torch.from_numpy(vel), draws=(u_pose, u_vel, r), **kw)
        for e in range(E):
            want_p, want_v, used = io.reset_noise(CH, pose[e], vel[e], u_pose[e], u_vel[e], r[e], **kw)
            assert used == (n_r if kw["radian"] != 0 else 0)
            np.testing.assert_allclose(got_p[e].numpy(), want_p, atol=1e-12)
            np.testing.assert_allclose(got_v[e].numpy(), want_v, atol=1e-12)
