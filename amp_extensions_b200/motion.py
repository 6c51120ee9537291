"""Reference motion clip, prepared on the host the way DeepMimicCore prepares it at load time.

Host-side (float64 numpy) counterpart of reference DeepMimicCore anim/Motion.cpp:356-442 (LoadJsonFrames,
PostProcessFrames), :170-191 (BuildFrameVel via cKinTree::CalcVel, anim/KinTree.cpp:1518-1556),
anim/KinController.cpp:144-175 (first-frame centring, cycle root delta).  The per-env interpolation at a
clip time (Motion.cpp:267-305, :498-527; KinTree.cpp:1577-1625) happens on the GPU from these tables.
"""
import json
import os

import numpy as np

from .character import JOINT_REVOLUTE, JOINT_ROOT, JOINT_SPHERICAL

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _quat_mul(a, b):
    """Hamilton product of (w, x, y, z) quaternions."""
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw])


def _quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def _normalize_angle(theta):
    """cMathUtil::NormalizeAngle (util/MathUtil.cpp:33-46)."""
    t = np.fmod(theta, 2 * np.pi)
    if t > np.pi:
        t -= 2 * np.pi
    elif t < -np.pi:
        t += 2 * np.pi
    return t


def _quat_to_axis_angle(q):
    """cMathUtil::QuaternionToAxisAngle (util/MathUtil.cpp:455-474)."""
    theta, axis = 0.0, np.array([0.0, 0.0, 1.0])
    q = np.array(q, dtype=np.float64)
    if q[0] > 1:
        q = q / np.linalg.norm(q)
    s2 = 1 - q[0] * q[0]
    sin_theta = np.sqrt(s2) if s2 >= 0 else np.nan
    if sin_theta > 0.000001:
        theta = _normalize_angle(2 * np.arccos(q[0]))
        axis = q[1:4] / sin_theta
    return axis, theta


class MotionClip:
    def __init__(self, frames, frame_times, frame_vels, loop_wrap, cycle_delta):
        self.frames = np.asarray(frames, dtype=np.float64)
        self.frame_times = np.asarray(frame_times, dtype=np.float64)
        self.frame_vels = np.asarray(frame_vels, dtype=np.float64)
        self.loop_wrap = bool(loop_wrap)
        self.cycle_delta = np.asarray(cycle_delta, dtype=np.float64)
        self.duration = float(self.frame_times[-1])  # Motion.cpp:444-449
        self.n_frames = self.frames.shape[0]

    @classmethod
    def from_raw(cls, raw, character, loop="wrap"):
        """raw: [n_frames, 1 + dof] rows of (duration, pose) as stored in a DeepMimic motion file."""
        raw = np.asarray(raw, dtype=np.float64)
        if raw.shape[1] != character.dof + 1:
            raise ValueError(f"DOF mismatch, char dof: {character.dof}, motion dof: {raw.shape[1] - 1}")
        durations = raw[:, 0]
        frames = raw[:, 1:].copy()
        n = frames.shape[0]
        # PostProcessFrames: cumulative start times, centre on the first frame (xz), normalise quaternions
        times = np.concatenate([[0.0], np.cumsum(durations)[:-1]])
        off = frames[0, 0:3].copy()
        off[1] = 0.0
        frames[:, 0:3] -= off
        for j in range(character.n_joints):
            if character.joint_type[j] in (JOINT_ROOT, JOINT_SPHERICAL):
                o = character.param_offset[j] + (3 if character.joint_type[j] == JOINT_ROOT else 0)
                frames[:, o:o + 4] /= np.linalg.norm(frames[:, o:o + 4], axis=1, keepdims=True)
        # KinController::PostProcessMotion: x/z relative to the first frame (a no-op after the centring above)
        frames[:, 0] -= frames[0, 0]
        frames[:, 2] -= frames[0, 2]
        vels = np.zeros_like(frames)
        for f in range(n - 1):
            vels[f] = cls._calc_vel(character, frames[f], frames[f + 1], times[f + 1] - times[f])
        if n > 1:
            vels[n - 1] = vels[n - 2]
        delta = frames[-1, 0:3] - frames[0, 0:3]
        delta[1] = 0.0
        return cls(frames, times, vels, loop == "wrap", delta)

    @classmethod
    def from_json(cls, path, character):
        with open(path) as f:
            d = json.load(f)
        return cls.from_raw(np.array(d["Frames"], dtype=np.float64), character, d.get("Loop", "none"))

    @classmethod
    def spinkick(cls, character):
        """humanoid3d_spinkick: the clip BASELINE.json's configs are quoted on (78 frames, wrap)."""
        z = np.load(os.path.join(DATA_DIR, "humanoid3d_spinkick.npz"))
        return cls.from_raw(z["frames_raw"], character, str(z["loop"]))

    @staticmethod
    def _calc_vel(ch, p0, p1, dt):
        """cKinTree::CalcVel (KinTree.cpp:1518-1556)."""
        v = np.zeros_like(p0)
        v[0:3] = (p1[0:3] - p0[0:3]) / dt
        axis, theta = _quat_to_axis_angle(_quat_mul(p1[3:7], _quat_conj(p0[3:7])))  # world frame, MathUtil.cpp:498
        v[3:6] = (theta / dt) * axis
        for j in range(1, ch.n_joints):
            o, sz = ch.param_offset[j], ch.param_size[j]
            if ch.joint_type[j] == JOINT_SPHERICAL:
                axis, theta = _quat_to_axis_angle(_quat_mul(_quat_conj(p0[o:o + 4]), p1[o:o + 4]))  # joint frame, :507
                v[o:o + 3] = (theta / dt) * axis
            elif ch.joint_type[j] == JOINT_REVOLUTE:
                v[o:o + sz] = (p1[o:o + sz] - p0[o:o + sz]) / dt
        return v
