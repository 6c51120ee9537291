"""CPU: known-answer tests that anchor the imitation-reward oracle (the reference C++ cannot be built
here, so these closed forms are the pin; see oracle/imitation_oracle.py header)."""
import os

import numpy as np
import pytest

from oracle import imitation_oracle as io

CH = io.HUMANOID3D
RAW = np.load(os.path.join(os.path.dirname(__file__), "..", "amp_extensions_b200", "data",
                           "humanoid3d_spinkick.npz"))["frames_raw"]


@pytest.fixture(scope="module")
def clip():
    return io.Clip(RAW, CH, "wrap")


def tpose():
    p = np.zeros(43)
    p[1] = 0.9
    offs, _ = io.param_layout(CH)
    p[3] = 1.0
    for j, t in enumerate(CH["joint_type"]):
        if t == io.SPHERICAL:
            p[offs[j]] = 1.0
    return p


def axis_angle_quat(axis, theta):
    axis = np.asarray(axis, dtype=np.float64) / np.linalg.norm(axis)
    return np.concatenate([[np.cos(theta / 2)], np.sin(theta / 2) * axis])


def test_param_layout_matches_survey():
    offs, sizes = io.param_layout(CH)
    assert offs == [0, 7, 11, 15, 19, 20, 24, 28, 29, 29, 33, 34, 38, 42, 43] and sum(sizes) == 43


def test_clip_tables(clip):
    assert clip.n == 78 and clip.duration == pytest.approx(1.283282, abs=1e-9)
    np.testing.assert_allclose(clip.cycle_delta, [-0.399868, 0.0, -0.275009], atol=2e-6)
    assert np.allclose(clip.frames[0, [0, 2]], 0.0)
    for k in (0, 5, 40, 76):
        np.testing.assert_allclose(clip.kin_pose(clip.times[k]), standardized(clip.frames[k]), atol=1e-12)
        wrapped = clip.kin_pose(2 * clip.duration + clip.times[k] + 1e-12)
        expect = standardized(clip.frames[k]).copy()
        expect[0:3] += 2 * clip.cycle_delta
        np.testing.assert_allclose(wrapped, expect, atol=1e-7)
    # frame velocity is the finite difference the loader takes (Motion.cpp:170-191)
    np.testing.assert_allclose(clip.vels[3, 0:3], (clip.frames[4, 0:3] - clip.frames[3, 0:3]) / 0.016666, rtol=1e-9)
    np.testing.assert_allclose(clip.vels[-1], clip.vels[-2])


def standardized(pose):
    p = pose.copy()
    if p[3] < 0:
        p[3:7] = -p[3:7]
    return p


def test_forward_kinematics_of_the_t_pose():
    p = tpose()
    at = np.asarray(CH["attach"])
    np.testing.assert_allclose(io.calc_joint_world_pos(CH, p, 0), [0, 0.9, 0])
    np.testing.assert_allclose(io.calc_joint_world_pos(CH, p, 8), p[0:3] + at[1] + at[6] + at[7] + at[8], atol=1e-12)
    np.testing.assert_allclose(io.calc_joint_world_pos(CH, p, 11), p[0:3] + at[9] + at[10] + at[11], atol=1e-12)
    # rotate the root by 90 deg about y: the right hip (attach +z) moves to +x ... R_y(90) maps z -> x
    p[3:7] = axis_angle_quat((0, 1, 0), np.pi / 2)
    np.testing.assert_allclose(io.calc_joint_world_pos(CH, p, 3), p[0:3] + [0.084887, 0, 0], atol=1e-12)
    # bend the right knee (revolute about local z) by 90 deg: the ankle offset (0,-l,0) becomes (l,0,0) in the hip frame
    p = tpose()
    p[19] = np.pi / 2
    np.testing.assert_allclose(io.calc_joint_world_pos(CH, p, 5), p[0:3] + at[3] + at[4] + [0.40987, 0, 0], atol=1e-12)


def test_heading_and_origin_transform():
    p = tpose()
    p[3:7] = axis_angle_quat((0, 1, 0), 0.7)
    assert io.calc_heading(p[3:7]) == pytest.approx(0.7)
    m = io.build_origin_trans(p)
    # a direction along the character's heading maps onto +x
    d = np.array([np.cos(0.7), 0, -np.sin(0.7), 0.0])
    np.testing.assert_allclose(m @ d, [1, 0, 0, 0], atol=1e-12)


def test_quat_theta_branches():
    assert io.quat_theta([1.0, 0, 0, 0]) == 0.0
    assert io.quat_theta(axis_angle_quat((1, 0, 0), 1e-5)) == 0.0          # sin(theta/2) <= 1e-4 -> 0
    assert io.quat_theta(axis_angle_quat((1, 0, 0), 0.5)) == pytest.approx(0.5)
    assert io.quat_theta(-axis_angle_quat((1, 0, 0), 0.5)) == pytest.approx(-0.5)  # antipodal: folded, squared later
    assert io.quat_theta(axis_angle_quat((0, 0, 1), 3.5)) == pytest.approx(3.5 - 2 * np.pi)


def test_reward_is_one_on_the_clip(clip):
    for t in (0.0, 0.31, 1.2, 1.283282 * 3 + 0.4):
        p, v = clip.kin_pose(t), clip.kin_vel(t)
        r, terms = io.calc_reward_imitate(CH, p, v, p, v, 0.0, True)
        assert r == pytest.approx(1.0, abs=1e-12) and np.allclose(terms, 1.0)
    # a kin origin offset moves both characters' reference frames consistently
    org = np.array([0.3, 0.2, -0.4])
    p1 = clip.kin_pose(0.5, org)
    p0 = clip.kin_pose(0.5)
    p0[0:3] += [org[0], 0.0, org[2]]   # the sim character stands on the real ground (y = 0)
    r, terms = io.calc_reward_imitate(CH, p0, clip.kin_vel(0.5), p1, clip.kin_vel(0.5), org[1], True)
    assert r == pytest.approx(1.0, abs=1e-12)


def test_single_joint_rotation_gives_the_closed_form_pose_error(clip):
    t = 0.45
    p1, v1 = clip.kin_pose(t), clip.kin_vel(t)
    theta = 0.3
    p0 = p1.copy()
    p0[7:11] = io.quat_mul(p1[7:11], axis_angle_quat((0.2, 1.0, -0.4), theta))   # chest, joint weight 0.5/4.8
    _, terms = io.calc_reward_imitate(CH, p0, v1, p1, v1, 0.0, True)
    assert terms[0] == pytest.approx(np.exp(-2.0 * (0.5 / 4.8) * theta ** 2), rel=1e-12)
    assert terms[1] == pytest.approx(1.0) and terms[3] == pytest.approx(1.0)
    # knee (revolute, weight 0.3/4.8) offset by 0.2 rad
    p0 = p1.copy()
    p0[19] += 0.2
    _, terms = io.calc_reward_imitate(CH, p0, v1, p1, v1, 0.0, True)
    assert terms[0] == pytest.approx(np.exp(-2.0 * (0.3 / 4.8) * 0.04), rel=1e-12)
    # velocity offsets: vel reward = exp(-0.1 * sum_j w_j |dv_j|^2), root term sees the root part only
    v0 = v1.copy()
    v0[11:14] += [0.5, -0.25, 1.0]   # neck angular velocity, weight 0.3/4.8
    _, terms = io.calc_reward_imitate(CH, p1, v0, p1, v1, 0.0, True)
    assert terms[1] == pytest.approx(np.exp(-0.1 * (0.3 / 4.8) * (0.25 + 0.0625 + 1.0)), rel=1e-12)
    assert terms[3] == pytest.approx(1.0)
    # root translation by d: root_err = |d|^2 -> exp(-5 |d|^2); end effectors are root relative in xz but not in y
    p0 = p1.copy()
    p0[0] += 0.1
    p0[2] -= 0.2
    _, terms = io.calc_reward_imitate(CH, p0, v1, p1, v1, 0.0, True)
    assert terms[3] == pytest.approx(np.exp(-5.0 * 0.05), rel=1e-12) and terms[2] == pytest.approx(1.0)


def integrate(pose, vel, dt):
    """Advance generalized coordinates by dt under generalized velocity vel (root: world frame; joints: local)."""
    offs, _ = io.param_layout(CH)
    p = pose.copy()
    p[0:3] += vel[0:3] * dt
    w = vel[3:6]
    n = np.linalg.norm(w)
    if n > 0:
        p[3:7] = io.quat_mul(axis_angle_quat(w, n * dt), pose[3:7])
    for j, t in enumerate(CH["joint_type"]):
        o = offs[j]
        if t == io.SPHERICAL:
            w = vel[o:o + 3]
            n = np.linalg.norm(w)
            if n > 0:
                p[o:o + 4] = io.quat_mul(pose[o:o + 4], axis_angle_quat(w, n * dt))
        elif t == io.REVOLUTE:
            p[o] += vel[o] * dt
    return p


def test_com_velocity_is_the_time_derivative_of_the_com(clip):
    rng = np.random.default_rng(0)
    p = clip.kin_pose(0.8)
    v = clip.kin_vel(0.8) + rng.normal(0, 0.5, 43)
    v[[6, 10, 14, 18, 23, 27, 32, 37, 41]] = 0.0  # 4th slot of angular velocities
    com, com_vel = io.calc_com(CH, p, v)
    h = 1e-6
    cp, _ = io.calc_com(CH, integrate(p, v, h), v)
    cm, _ = io.calc_com(CH, integrate(p, v, -h), v)
    np.testing.assert_allclose(com_vel, (cp - cm) / (2 * h), atol=1e-7)
    # total mass 45: COM of the T-pose is the mass-weighted mean of the body centres
    tp = tpose()
    com, _ = io.calc_com(CH, tp, np.zeros(43))
    centres = np.array([io.calc_joint_world_pos(CH, tp, j) + np.asarray(CH["body_attach"][j]) for j in range(15)])
    np.testing.assert_allclose(com, (np.asarray(CH["body_mass"])[:, None] * centres).sum(0) / 45.0, atol=1e-12)


def test_product_clip_tables_equal_the_oracle_tables(clip):
    from amp_extensions_b200.character import humanoid3d
    from amp_extensions_b200.motion import MotionClip
    ch = humanoid3d()
    mc = MotionClip.spinkick(ch)
    np.testing.assert_allclose(mc.frames, clip.frames, atol=1e-14)
    np.testing.assert_allclose(mc.frame_times, clip.times, atol=1e-14)
    np.testing.assert_allclose(mc.frame_vels, clip.vels, atol=1e-10)
    np.testing.assert_allclose(mc.cycle_delta, clip.cycle_delta, atol=1e-14)
    assert mc.duration == clip.duration and mc.loop_wrap
    np.testing.assert_allclose(ch.joint_weights(), np.asarray(CH["diff_weight"]) / 4.8)


# ---------------------------------------------------------------------------------------------------
# record_state (CtController.cpp:378-495): closed-form answers


def _tpose():
    pose = np.zeros(43)
    pose[1] = 0.9
    for j, jt in enumerate(io.HUMANOID3D["joint_type"]):
        if jt in (io.ROOT, io.SPHERICAL):
            o = io.param_layout(io.HUMANOID3D)[0][j] + (3 if jt == io.ROOT else 0)
            pose[o] = 1.0
    return pose


def test_record_state_of_the_t_pose_is_the_skeleton_table():
    """Identity rotations, zero velocity: body i sits at sum of attach offsets along its chain + its body offset,
    every normal is +y, every tangent +x, all velocities are 0, state[0] is the root height."""
    ch = io.HUMANOID3D
    st = io.record_state(ch, _tpose(), np.zeros(43))
    assert st.shape == (226,) and st[0] == 0.9
    for i in range(15):
        chain, j = np.zeros(3), i
        while j > 0:
            chain += np.asarray(ch["attach"][j])
            j = ch["parent"][j]
        expect = chain + np.asarray(ch["body_attach"][i])
        np.testing.assert_allclose(st[1 + 9 * i: 4 + 9 * i], expect, atol=1e-12)
        np.testing.assert_allclose(st[4 + 9 * i: 7 + 9 * i], [0, 1, 0], atol=1e-12)
        np.testing.assert_allclose(st[7 + 9 * i: 10 + 9 * i], [1, 0, 0], atol=1e-12)
    assert not st[136:].any()


def test_record_state_is_heading_and_translation_invariant():
    """Turning the whole character about the vertical axis and moving it in the plane changes nothing but the
    root's own rotation / velocity features (RecordWorldRootRot = true keeps those in the world frame)."""
    ch = io.HUMANOID3D
    clip = io.Clip(RAW, ch, "wrap")
    pose, vel = clip.kin_pose(0.37), clip.kin_vel(0.37)
    a = 0.8
    qy = np.array([np.cos(a / 2), 0, np.sin(a / 2), 0])
    R = io.rotate_mat_quat(qy)[0:3, 0:3]
    pose2, vel2 = pose.copy(), vel.copy()
    pose2[0:3] = R @ pose[0:3] + np.array([1.5, 0, -2.0])
    pose2[3:7] = io.quat_mul(qy, pose[3:7])
    vel2[0:3], vel2[3:6] = R @ vel[0:3], R @ vel[3:6]
    s1, s2 = io.record_state(ch, pose, vel), io.record_state(ch, pose2, vel2)
    keep = np.ones(226, bool)
    keep[4:10] = False        # root normal / tangent (world)
    keep[136:142] = False     # root linear / angular velocity (world)
    np.testing.assert_allclose(s1[keep], s2[keep], atol=1e-9)
    np.testing.assert_allclose(s2[4:7], R @ s1[4:7], atol=1e-9)
    np.testing.assert_allclose(s2[136:139], R @ s1[136:139], atol=1e-9)
    # with every Record* flag off the root features are in the heading frame too: fully invariant
    f = dict(record_world_root_rot=False)
    np.testing.assert_allclose(io.record_state(ch, pose, vel, **f), io.record_state(ch, pose2, vel2, **f), atol=1e-9)


def test_record_state_velocities_are_time_derivatives_of_positions():
    """Body linear velocities equal d/dt of the body positions along the clip (world frame, RecordAllWorld)."""
    ch = io.HUMANOID3D
    clip = io.Clip(RAW, ch, "wrap")
    t, dt = 0.401, 1e-5       # inside one frame interval: the interpolated pose is smooth there
    p0, p1 = clip.kin_pose(t - dt), clip.kin_pose(t + dt)
    vel = io.calc_vel(ch, p0, p1, 2 * dt)   # the pose path's own generalized velocity
    s0 = io.record_state(ch, p0, vel, record_all_world=True)
    s1 = io.record_state(ch, p1, vel, record_all_world=True)
    sm = io.record_state(ch, clip.kin_pose(t), vel, record_all_world=True)
    for i in range(15):
        fd = (s1[1 + 9 * i: 4 + 9 * i] - s0[1 + 9 * i: 4 + 9 * i]) / (2 * dt)
        np.testing.assert_allclose(sm[136 + 6 * i: 139 + 6 * i], fd, atol=2e-4)
