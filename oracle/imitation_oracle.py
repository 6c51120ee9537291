"""CPU oracle for the DeepMimic imitation reward — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy float64 restatement, kept deliberately literal (4x4 homogeneous matrices, 6-D spatial vectors, the
Jacobian-based centre-of-mass velocity), of the reference C++ under
deepmimic/deepmimic/DeepMimicCore/ (cited per function as FILE:LINE).

PARITY PINNED (kinematics), PARTLY RESTATED (glue): oracle/ref_build.py compiles the reference's own
util/MathUtil.cpp, anim/KinTree.cpp, anim/Motion.cpp, sim/RBDUtil.cpp, sim/SpAlg.cpp, sim/RBDModel.cpp, util/JsonUtil.cpp,
util/FileUtil.cpp and util/json/*.cpp where they lie into oracle/_ref/libdmref.so.  Eigen 3.3.7 is absent from the
image (as are Bullet 2.88 and SWIG), so those files are compiled against the stand-in headers in oracle/eigen_shim;
the ~60 lines of cSceneImitate::CalcRewardImitate that combine the error terms, the kinematic character's loop /
origin handling and CtController's state layout need Bullet-dependent classes and are restated in oracle/ref_driver.cpp
around calls into the compiled reference functions.  tests/test_imitation_ref.py checks every function of this file
against that library (live, and through the committed vectors tests/golden/imitation_ref_golden.npz written by
tests/golden/make_imitation_ref_golden.py): clip tables, clip sampling, per-joint pose / velocity errors, forward
kinematics, heading and origin transform, centre of mass and its velocity, state features, reward and its five terms
agree to <= 1e-9 (observed 1e-14).  The closed-form known answers of tests/test_imitation_oracle.py remain as an
independent anchor: reward == 1 at pose == clip(t); a chest rotation by theta gives pose_err = w_chest*theta^2; FK
against hand-computed joint positions; COM velocity against a finite difference of the COM position.

One deliberate deviation, shared with the CUDA path and stated in DESIGN.md: the simulated character's COM velocity comes from the same kinematic formula as the kinematic
character's (RBDUtil.cpp:572-613); the reference takes it from Bullet body velocities
(SimCharacter.cpp:398-436), which do not exist for a learned-dynamics state.

Eigen arithmetic that is not under the reference tree (Quaternion::slerp, q*v) is restated from Eigen
3.3.7 (pinned in the reference's setup.md:27) - here and, independently, in oracle/eigen_shim/Eigen/Core.
"""
import numpy as np

# ---- humanoid3d character (deepmimic/deepmimic/data/characters/humanoid3d.txt) ------------------------
ROOT, SPHERICAL, REVOLUTE, FIXED = 0, 1, 2, 3
HUMANOID3D = dict(
    joint_type=[ROOT, SPHERICAL, SPHERICAL, SPHERICAL, REVOLUTE, SPHERICAL, SPHERICAL, REVOLUTE, FIXED, SPHERICAL,
                REVOLUTE, SPHERICAL, SPHERICAL, REVOLUTE, FIXED],
    parent=[-1, 0, 1, 0, 3, 4, 1, 6, 7, 0, 9, 10, 1, 12, 13],
    attach=[(0, 0, 0), (0, 0.236151, 0), (0, 0.223894, 0), (0, 0, 0.084887), (0, -0.421546, 0), (0, -0.40987, 0),
            (-0.02405, 0.2435, 0.18311), (0, -0.274788, 0), (0, -0.258947, 0), (0, 0, -0.084887), (0, -0.421546, 0),
            (0, -0.40987, 0), (-0.02405, 0.2435, -0.18311), (0, -0.274788, 0), (0, -0.258947, 0)],
    is_end_eff=[0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1],
    diff_weight=[1, 0.5, 0.3, 0.5, 0.3, 0.2, 0.3, 0.2, 0, 0.5, 0.3, 0.2, 0.3, 0.2, 0],
    body_mass=[6.0, 14.0, 2.0, 4.5, 3.0, 1.0, 1.5, 1.0, 0.5, 4.5, 3.0, 1.0, 1.5, 1.0, 0.5],
    body_attach=[(0, 0.07, 0), (0, 0.12, 0), (0, 0.175, 0), (0, -0.21, 0), (0, -0.2, 0), (0.045, -0.0225, 0),
                 (0, -0.14, 0), (0, -0.12, 0), (0, 0, 0), (0, -0.21, 0), (0, -0.2, 0), (0.045, -0.0225, 0),
                 (0, -0.14, 0), (0, -0.12, 0), (0, 0, 0)],
)
_PARAM_SIZE = {ROOT: 7, SPHERICAL: 4, REVOLUTE: 1, FIXED: 0}


def param_layout(ch):
    """KinTree.cpp:824-850 (sizes), :1053-1062 (offsets)."""
    sizes = [_PARAM_SIZE[t] for t in ch["joint_type"]]
    offs = list(np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(int))
    return offs, sizes


# ---- util/MathUtil.cpp --------------------------------------------------------------------------------

def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_diff(q0, q1):
    """MathUtil.cpp:527-530."""
    return quat_mul(q1, quat_conj(q0))


def normalize_angle(theta):
    """MathUtil.cpp:33-46."""
    t = np.fmod(theta, 2 * np.pi)
    if t > np.pi:
        t = -2 * np.pi + t
    elif t < -np.pi:
        t = 2 * np.pi + t
    return t


def quat_theta(dq):
    """MathUtil.cpp:538-554."""
    theta = 0.0
    q1 = np.array(dq, dtype=np.float64)
    if q1[0] > 1:
        q1 = q1 / np.linalg.norm(q1)
    with np.errstate(invalid="ignore"):
        sin_theta = np.sqrt(1 - q1[0] * q1[0])
    if sin_theta > 0.0001:
        theta = normalize_angle(2 * np.arccos(q1[0]))
    return theta


def quat_to_axis_angle(q):
    """MathUtil.cpp:455-474."""
    theta, axis = 0.0, np.array([0.0, 0.0, 1.0])
    q1 = np.array(q, dtype=np.float64)
    if q1[0] > 1:
        q1 = q1 / np.linalg.norm(q1)
    with np.errstate(invalid="ignore"):
        sin_theta = np.sqrt(1 - q1[0] * q1[0])
    if sin_theta > 0.000001:
        theta = normalize_angle(2 * np.arccos(q1[0]))
        axis = q1[1:4] / sin_theta
    return axis, theta


def quat_rot_vec(q, v):
    """MathUtil.cpp:561-566 = Eigen 3.3.7 Quaternion::_transformVector (no normalisation)."""
    u = np.asarray(q[1:4], dtype=np.float64)
    uv = 2.0 * np.cross(u, v)
    return np.asarray(v, dtype=np.float64) + q[0] * uv + np.cross(u, uv)


def rotate_mat_quat(q):
    """MathUtil.cpp:212-242."""
    w, x, y, z = q
    m = np.eye(4)
    sqw, sqx, sqy, sqz = w * w, x * x, y * y, z * z
    invs = 1 / (sqx + sqy + sqz + sqw)
    m[0, 0] = (sqx - sqy - sqz + sqw) * invs
    m[1, 1] = (-sqx + sqy - sqz + sqw) * invs
    m[2, 2] = (-sqx - sqy + sqz + sqw) * invs
    t1, t2 = x * y, z * w
    m[1, 0] = 2.0 * (t1 + t2) * invs
    m[0, 1] = 2.0 * (t1 - t2) * invs
    t1, t2 = x * z, y * w
    m[2, 0] = 2.0 * (t1 - t2) * invs
    m[0, 2] = 2.0 * (t1 + t2) * invs
    t1, t2 = y * z, x * w
    m[2, 1] = 2.0 * (t1 + t2) * invs
    m[1, 2] = 2.0 * (t1 - t2) * invs
    return m


def rotate_mat_axis(axis, theta):
    """MathUtil.cpp:193-210."""
    c, s = np.cos(theta), np.sin(theta)
    x, y, z = axis
    return np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s, 0],
                     [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s, 0],
                     [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c), 0],
                     [0, 0, 0, 1]])


def translate_mat(t):
    m = np.eye(4)
    m[0:3, 3] = t[0:3]
    return m


def slerp(a, b, t):
    """Eigen 3.3.7 QuaternionBase::slerp (call sites KinTree.cpp:1595, :1612)."""
    one = 1.0 - np.finfo(np.float64).eps
    d = float(np.dot(a, b))
    ad = abs(d)
    if ad >= one:
        s0, s1 = 1.0 - t, t
    else:
        theta = np.arccos(ad)
        st = np.sin(theta)
        s0, s1 = np.sin((1.0 - t) * theta) / st, np.sin(t * theta) / st
    if d < 0:
        s1 = -s1
    return s0 * np.asarray(a) + s1 * np.asarray(b)


# ---- anim/KinTree.cpp ---------------------------------------------------------------------------------

def child_parent_trans(ch, pose, j):
    """KinTree.cpp:1081-1116, :1806-1878; attach rotations are zero for humanoid3d."""
    offs, _ = param_layout(ch)
    t = ch["joint_type"][j]
    A = np.eye(4)
    if t != ROOT:  # the root's attach point is zeroed at load, KinTree.cpp:1064-1067
        A[0:3, 3] = ch["attach"][j]
    if t == ROOT:
        return A @ translate_mat(pose[0:3]) @ rotate_mat_quat(pose[3:7])
    if t == REVOLUTE:
        return A @ rotate_mat_axis((0, 0, 1), pose[offs[j]])
    if t == SPHERICAL:
        return A @ rotate_mat_quat(pose[offs[j]:offs[j] + 4])
    return A


def joint_world_trans(ch, pose, j):
    """KinTree.cpp:1126-1139."""
    m = np.eye(4)
    cur = j
    while cur != -1:
        m = child_parent_trans(ch, pose, cur) @ m
        cur = ch["parent"][cur]
    return m


def calc_joint_world_pos(ch, pose, j):
    """KinTree.cpp:586-601."""
    return (joint_world_trans(ch, pose, j) @ np.array([0, 0, 0, 1.0]))[0:3]


def calc_heading(q):
    """KinTree.cpp:1667-1675."""
    d = quat_rot_vec(q, np.array([1.0, 0, 0]))
    return np.arctan2(-d[2], d[0])


def build_origin_trans(pose):
    """KinTree.cpp:1693-1712."""
    origin = np.array(pose[0:3], dtype=np.float64)
    origin[1] = 0
    rot = rotate_mat_axis((0, 1, 0), -calc_heading(pose[3:7]))
    return rot @ translate_mat(-origin)


def calc_pose_err(ch, j, pose0, pose1):
    """KinTree.cpp:1367-1398 for a non-root joint (+ CalcJointPoseDiff :1477-1509)."""
    offs, sizes = param_layout(ch)
    o, sz = offs[j], sizes[j]
    if ch["joint_type"][j] == SPHERICAL:
        th = quat_theta(quat_diff(pose0[o:o + 4], pose1[o:o + 4]))
        return th * th
    d = pose1[o:o + sz] - pose0[o:o + sz]
    return float(np.dot(d, d))


def calc_vel_err(ch, j, vel0, vel1):
    """KinTree.cpp:1445-1462, :1511-1516."""
    offs, sizes = param_layout(ch)
    d = vel1[offs[j]:offs[j] + sizes[j]] - vel0[offs[j]:offs[j] + sizes[j]]
    return float(np.dot(d, d))


def calc_vel(ch, p0, p1, dt):
    """KinTree.cpp:1518-1556."""
    offs, sizes = param_layout(ch)
    v = np.zeros_like(p0)
    v[0:3] = (p1[0:3] - p0[0:3]) / dt
    axis, theta = quat_to_axis_angle(quat_diff(p0[3:7], p1[3:7]))  # MathUtil.cpp:498-505
    v[3:6] = (theta / dt) * axis
    for j in range(1, len(offs)):
        o, sz = offs[j], sizes[j]
        if ch["joint_type"][j] == SPHERICAL:
            axis, theta = quat_to_axis_angle(quat_mul(quat_conj(p0[o:o + 4]), p1[o:o + 4]))  # MathUtil.cpp:507-515
            v[o:o + 3] = (theta / dt) * axis
        else:
            v[o:o + sz] = (p1[o:o + sz] - p0[o:o + sz]) / dt
    return v


def lerp_poses(ch, p0, p1, lerp):
    """KinTree.cpp:1577-1620."""
    offs, sizes = param_layout(ch)
    out = np.zeros_like(p0)
    out[0:3] = (1 - lerp) * p0[0:3] + lerp * p1[0:3]
    r = slerp(p0[3:7], p1[3:7], lerp)
    out[3:7] = r / np.linalg.norm(r)
    for j in range(1, len(offs)):
        o, sz = offs[j], sizes[j]
        if ch["joint_type"][j] == SPHERICAL:
            out[o:o + 4] = slerp(p0[o:o + 4], p1[o:o + 4], lerp)
        else:
            out[o:o + sz] = (1 - lerp) * p0[o:o + sz] + lerp * p1[o:o + sz]
    return out


# ---- anim/Motion.cpp, KinController.cpp, MotionController.cpp, KinCharacter.cpp ------------------------

class Clip:
    def __init__(self, raw, ch, loop="wrap"):
        """Motion.cpp:356-442 (load + PostProcessFrames), :170-191 (frame velocities),
        KinController.cpp:144-175 (centring, cycle delta)."""
        raw = np.asarray(raw, dtype=np.float64)
        offs, _ = param_layout(ch)
        self.ch = ch
        self.loop = loop == "wrap"
        dur = raw[:, 0]
        frames = raw[:, 1:].copy()
        n = frames.shape[0]
        times = np.zeros(n)
        cur = 0.0
        root_off = frames[0, 0:3].copy()
        root_off[1] = 0
        for f in range(n):
            times[f] = cur
            cur += dur[f]
            frames[f, 0:3] -= root_off
            frames[f, 3:7] /= np.linalg.norm(frames[f, 3:7])  # KinTree.cpp:1558-1575
            for j in range(1, len(offs)):
                if ch["joint_type"][j] == SPHERICAL:
                    frames[f, offs[j]:offs[j] + 4] /= np.linalg.norm(frames[f, offs[j]:offs[j] + 4])
        beg = frames[0, 0:3].copy()
        for f in range(n):  # KinController.cpp:144-160
            frames[f, 0] -= beg[0]
            frames[f, 2] -= beg[2]
        self.frames, self.times, self.n = frames, times, n
        self.duration = times[-1]  # Motion.cpp:444-449
        self.vels = np.zeros_like(frames)
        for f in range(n - 1):
            self.vels[f] = calc_vel(ch, frames[f], frames[f + 1], times[f + 1] - times[f])
        self.vels[n - 1] = self.vels[n - 2]
        self.cycle_delta = frames[-1, 0:3] - frames[0, 0:3]  # KinController.cpp:162-175
        self.cycle_delta[1] = 0

    def cycle_count(self, time):
        """Motion.cpp:488-496."""
        c = int(np.floor(time / self.duration))
        return c if self.loop else min(max(c, 0), 1)

    def index_blend(self, time):
        """Motion.cpp:498-527."""
        if not self.loop:
            if time <= 0:
                return 0, 0.0
            if time >= self.duration:
                return self.n - 2, 1.0
        time = time - self.cycle_count(time) * self.duration
        idx = int(np.searchsorted(self.times, time, side="right")) - 1
        t0, t1 = self.times[idx], self.times[idx + 1]
        return idx, (time - t0) / (t1 - t0)

    def kin_pose(self, time, origin=(0.0, 0.0, 0.0)):
        """cKinCharacter::CalcPose (KinCharacter.cpp:573-598) with mOriginRot = identity."""
        idx, blend = self.index_blend(time)
        blend = min(max(blend, 0.0), 1.0)  # Motion.cpp:252
        pose = lerp_poses(self.ch, self.frames[idx], self.frames[idx + 1], blend)  # Motion.cpp:267-274
        if self.loop:  # MotionController.cpp:25-41, :144-153
            pose[0:3] += self.cycle_count(time) * self.cycle_delta
        if pose[3] < 0:  # StandardizeQuat, MathUtil.cpp:48-59
            pose[3:7] = -pose[3:7]
        pose[0:3] += np.asarray(origin, dtype=np.float64)
        return pose

    def kin_vel(self, time):
        """cKinCharacter::CalcVel (KinCharacter.cpp:622-640) -> Motion.cpp:276-305."""
        if not self.loop and time >= self.duration:
            return np.zeros(self.frames.shape[1])
        idx, blend = self.index_blend(time)
        return (1.0 - blend) * self.vels[idx] + blend * self.vels[idx + 1]  # KinTree.cpp:1622-1625


# ---- sim/SpAlg.cpp (spatial transforms as (E, r)) ------------------------------------------------------

def sp_mat_to_trans(m):
    """SpAlg.cpp:151-159."""
    E = m[0:3, 0:3].copy()
    r = -E.T @ m[0:3, 3]
    return E, r


def sp_inv_trans(X):
    E, r = X
    return E.T.copy(), -E @ r


def sp_comp_trans(X0, X1):
    """SpAlg.cpp:334-343."""
    E0, r0 = X0
    E1, r1 = X1
    return E0 @ E1, r1 + E1.T @ r0


def sp_apply_trans_m(X, sv):
    """SpAlg.cpp:230-242."""
    E, r = X
    o0, v0 = sv[0:3], sv[3:6]
    return np.concatenate([E @ o0, E @ (v0 - np.cross(r, o0))])


def sp_apply_inv_trans_m(X, sv):
    """SpAlg.cpp:282-294."""
    E, r = X
    o0, v0 = sv[0:3], sv[3:6]
    return np.concatenate([E.T @ o0, E.T @ v0 + np.cross(r, E.T @ o0)])


# ---- sim/RBDUtil.cpp -----------------------------------------------------------------------------------

def joint_subspace(ch, pose, j):
    """RBDUtil.cpp:798-893."""
    offs, sizes = param_layout(ch)
    t = ch["joint_type"][j]
    S = np.zeros((6, sizes[j]))
    if t == ROOT:
        E = rotate_mat_quat(pose[3:7])[0:3, 0:3]
        S[3:6, 0:3] = E.T
        S[0:3, 3:6] = E.T
    elif t == REVOLUTE:
        S[2, 0] = 1
    elif t == SPHERICAL:
        S[0, 0] = S[1, 1] = S[2, 2] = 1
    return S


def end_effector_jacobian(ch, pose, joint_id):
    """RBDUtil.cpp:225-249."""
    offs, sizes = param_layout(ch)
    J = np.zeros((6, len(pose)))
    cur = joint_id
    trans = (np.eye(3), np.zeros(3))
    while cur != -1:
        S = joint_subspace(ch, pose, cur)
        for col in range(S.shape[1]):
            J[:, offs[cur] + col] = sp_apply_trans_m(trans, S[:, col])
        parent_child = sp_inv_trans(sp_mat_to_trans(child_parent_trans(ch, pose, cur)))  # RBDUtil.cpp:777-790
        trans = sp_comp_trans(trans, parent_child)
        cur = ch["parent"][cur]
    for col in range(J.shape[1]):
        J[:, col] = sp_apply_inv_trans_m(trans, J[:, col])
    return J


def calc_com(ch, pose, vel):
    """RBDUtil.cpp:572-613 -> (com, com_vel)."""
    com, com_vel, total = np.zeros(3), np.zeros(3), 0.0
    for j in range(len(ch["joint_type"])):
        m = ch["body_mass"][j]
        body_joint = translate_mat(np.asarray(ch["body_attach"][j], dtype=np.float64))  # KinTree.cpp:1156-1166
        world = joint_world_trans(ch, pose, j) @ body_joint
        world_com = (world @ np.array([0, 0, 0, 1.0]))[0:3]
        sv = end_effector_jacobian(ch, pose, j) @ vel  # RBDUtil.cpp:490-496
        sv = sp_apply_trans_m((np.eye(3), world_com), sv)
        com += m * world_com
        com_vel += m * sv[3:6]
        total += m
    return com / total, com_vel / total


# ---- sim/CtController.cpp: the env's state vector ---------------------------------------------------------

def record_state(ch, pose, vel, record_all_world=False, record_world_root_pos=False, record_world_root_rot=True,
                 vel_scale=1.0):
    """cCtController::BuildStatePose / BuildStateVel (CtController.cpp:378-495) for a character whose body parts
    follow the pose kinematically (cKinTree::BodyWorldTrans, KinTree.cpp:1148-1166; what the simulator holds right
    after SetPose / SetVel), plane ground at y = 0.  Literal: 4x4 transforms for positions and rotations, the
    spatial Jacobian (RBDUtil.cpp:225-249, :490-496) for the body velocities.  Flag defaults are those of
    data/controllers/humanoid3d_rot_ctrl.txt."""
    pose, vel = np.asarray(pose, dtype=np.float64), np.asarray(vel, dtype=np.float64)
    nj = len(ch["joint_type"])
    origin_trans = build_origin_trans(pose)
    origin_rot = origin_trans[0:3, 0:3]
    root_pos = np.array([pose[0], pose[1], pose[2], 1.0])
    root_pos_rel = origin_trans @ root_pos
    out_pose = np.zeros(1 + 9 * nj)
    out_vel = np.zeros(6 * nj)
    out_pose[0] = root_pos_rel[1]
    for i in range(nj):
        body_joint = translate_mat(np.asarray(ch["body_attach"][i], dtype=np.float64))
        world = joint_world_trans(ch, pose, i) @ body_joint
        pos = world @ np.array([0, 0, 0, 1.0])
        rot = world[0:3, 0:3]
        if not record_all_world and (not record_world_root_pos or i != 0):
            pos = origin_trans @ pos - root_pos_rel
        if not record_all_world and (not record_world_root_rot or i != 0):
            rot = origin_rot @ rot
        out_pose[1 + 9 * i: 4 + 9 * i] = pos[0:3]
        out_pose[4 + 9 * i: 7 + 9 * i] = rot @ np.array([0, 1.0, 0])   # CalcNormalTangent, MathUtil.cpp:622-628
        out_pose[7 + 9 * i: 10 + 9 * i] = rot @ np.array([1.0, 0, 0])
        world_com = (world @ np.array([0, 0, 0, 1.0]))[0:3]
        sv = end_effector_jacobian(ch, pose, i) @ vel
        sv = sp_apply_trans_m((np.eye(3), world_com), sv)
        ang, lin = sv[0:3], sv[3:6]
        if not record_all_world and (not record_world_root_rot or i != 0):
            ang, lin = origin_rot @ ang, origin_rot @ lin
        out_vel[6 * i: 6 * i + 3] = lin * vel_scale
        out_vel[6 * i + 3: 6 * i + 6] = ang * vel_scale
    return np.concatenate([out_pose, out_vel])


# ---- scenes/SceneImitate.cpp ---------------------------------------------------------------------------

def calc_reward_imitate(ch, pose0, vel0, pose1, vel1, ground_h1=0.0, return_terms=False):
    """SceneImitate.cpp:7-127 on a plane ground (h = 0 under the simulated character)."""
    pose_w, vel_w, end_eff_w, root_w, com_w = 0.5, 0.05, 0.15, 0.2, 0.1
    total_w = pose_w + vel_w + end_eff_w + root_w + com_w
    pose_w, vel_w, end_eff_w, root_w, com_w = (w / total_w for w in (pose_w, vel_w, end_eff_w, root_w, com_w))
    nj = len(ch["joint_type"])
    jw = np.asarray(ch["diff_weight"], dtype=np.float64)
    jw = jw / np.abs(jw).sum()  # SceneImitate.cpp:300-312
    pose_scale, vel_scale = 2.0 / 15 * nj, 0.1 / 15 * nj
    end_eff_scale, root_scale, com_scale, err_scale = 10.0, 5.0, 10.0, 1.0
    origin_trans, kin_origin_trans = build_origin_trans(pose0), build_origin_trans(pose1)
    _, com_vel0 = calc_com(ch, pose0, vel0)
    _, com_vel1 = calc_com(ch, pose1, vel1)
    root_pos0, root_pos1 = np.array(pose0[0:3]), np.array(pose1[0:3])
    th = quat_theta(quat_diff(pose0[3:7], pose1[3:7]))
    pose_err = jw[0] * th * th                                   # :67, KinTree.cpp:1406-1411
    d = vel1[3:7] - vel0[3:7]
    vel_err = jw[0] * float(np.dot(d, d))                        # :68, KinTree.cpp:1470-1474
    end_eff_err = 0.0
    for j in range(1, nj):
        pose_err += jw[j] * calc_pose_err(ch, j, pose0, pose1)
        vel_err += jw[j] * calc_vel_err(ch, j, vel0, vel1)
        if ch["is_end_eff"][j]:
            pos0 = calc_joint_world_pos(ch, pose0, j)
            pos1 = calc_joint_world_pos(ch, pose1, j)
            rel0 = np.append(pos0 - root_pos0, 0.0)
            rel1 = np.append(pos1 - root_pos1, 0.0)
            rel0[1] = pos0[1] - 0.0
            rel1[1] = pos1[1] - ground_h1
            rel0 = origin_trans @ rel0
            rel1 = kin_origin_trans @ rel1
            end_eff_err += float(np.dot(rel1 - rel0, rel1 - rel0))
    root_pos1 = root_pos1.copy()
    root_pos1[1] -= ground_h1
    root_pos_err = float(np.dot(root_pos0 - root_pos1, root_pos0 - root_pos1))
    root_rot_err = th * th
    dv = vel1[0:3] - vel0[0:3]
    root_err = root_pos_err + 0.1 * root_rot_err + 0.01 * float(np.dot(dv, dv)) + 0.001 * float(np.dot(d, d))
    dc = com_vel1 - com_vel0
    com_err = 0.1 * float(np.dot(dc, dc))
    terms = np.array([np.exp(-err_scale * pose_scale * pose_err), np.exp(-err_scale * vel_scale * vel_err),
                      np.exp(-err_scale * end_eff_scale * end_eff_err), np.exp(-err_scale * root_scale * root_err),
                      np.exp(-err_scale * com_scale * com_err)])
    reward = float(np.dot([pose_w, vel_w, end_eff_w, root_w, com_w], terms))
    return (reward, terms) if return_terms else reward


def imitation_reward_batch(ch, clip, pose, vel, kin_time, kin_origin=None):
    """Per-env loop over calc_reward_imitate against clip(t); returns (reward[E], terms[E,5])."""
    E = pose.shape[0]
    rew, terms = np.zeros(E), np.zeros((E, 5))
    for e in range(E):
        org = (0.0, 0.0, 0.0) if kin_origin is None else kin_origin[e]
        p1 = clip.kin_pose(float(kin_time[e]), org)
        v1 = clip.kin_vel(float(kin_time[e]))
        rew[e], terms[e] = calc_reward_imitate(ch, np.asarray(pose[e], dtype=np.float64),
                                               np.asarray(vel[e], dtype=np.float64), p1, v1, org[1], True)
    return rew, terms


# ---- reset noise (anim/KinCharacter.cpp:340-532) ---------------------------------------------------------

def euler_to_quat(euler):
    """cMathUtil::EulerToQuaternion (MathUtil.cpp:423-429) over EulerToAxisAngle (:347-378) and
    AxisAngleToQuaternion (:455-466)."""
    x, y, z = float(euler[0]), float(euler[1]), float(euler[2])
    xs, xc, ys, yc, zs, zc = np.sin(x), np.cos(x), np.sin(y), np.cos(y), np.sin(z), np.cos(z)
    c = (yc * zc + xs * ys * zs + xc * zc + xc * yc - 1) * 0.5
    c = min(max(c, -1.0), 1.0)
    theta = np.arccos(c)
    if abs(theta) < 0.00001:
        axis = np.array([0.0, 0.0, 1.0])
    else:
        m21 = xs * yc - xc * ys * zs + xs * zc
        m02 = xc * ys * zc + xs * zs + ys
        m10 = yc * zs - xs * ys * zc + xc * zs
        axis = np.array([m21, m02, m10]) / np.sqrt(m21 * m21 + m02 * m02 + m10 * m10)
    h = theta / 2
    return np.array([np.cos(h), np.sin(h) * axis[0], np.sin(h) * axis[1], np.sin(h) * axis[2]])


def reset_noise(ch, pose, vel, u_pose, u_vel, r, noise_bef_rot=False, noise_min=0.0, noise_max=0.0, radian=0.0,
                rot_vel_w_pose=False, vel_noise=False, interp=1.0, knee_rot=False):
    """cKinCharacter::AddNoise (KinCharacter.cpp:340-352) = AddNoisePoseVel (:354-366) and RandomRotatePoseVel
    (:367-532, root through cCharacter::RotateRoot, Character.cpp:210-216) in the order `noise_bef_rot` selects, for ONE
    character.  The reference's random draws are inputs: u_pose / u_vel [dof] in [0, 1) become U(noise_min, noise_max);
    r in [-1, 1) times `radian` are RandomRotatePoseVel's draws in the order it makes them (root yaw; 1 per noisy
    revolute joint, 3 per noisy spherical joint; with vel_noise 3 for the root's angular velocity, then again per
    joint).  The joint-index rules are the reference's, quirks included: knees (4, 10) only with knee_rot, hips and
    ankles (3, 5, 9, 11) never, and the revolute branch of the velocity noise tests `!(j == 4 || j != 10)` (:506), which
    is true for joint 10 alone.  Returns (pose, vel, number of r values consumed)."""
    offs, sizes = param_layout(ch)
    pose = np.array(pose, dtype=np.float64)
    vel = np.array(vel, dtype=np.float64)
    used = [0]

    def add_noise_pose_vel():
        if noise_min == 0 and noise_max == 0:
            return
        pose[:] = pose + (noise_min + (noise_max - noise_min) * np.asarray(u_pose, dtype=np.float64))
        vel[:] = vel + (noise_min + (noise_max - noise_min) * np.asarray(u_vel, dtype=np.float64))

    def draw():
        v = radian * float(r[used[0]])
        used[0] += 1
        return v

    def random_rotate():
        if radian == 0:
            return
        yaw = draw()
        rot = np.array([np.cos(yaw / 2), 0.0, np.sin(yaw / 2), 0.0])
        q = quat_mul(rot, pose[3:7])
        pose[3:7] = q / np.linalg.norm(q)
        vel[:] = interp * vel            # root velocity, root angular velocity and every joint's segment (:398-417)
        for j in range(1, len(sizes)):
            o, t = offs[j], ch["joint_type"][j]
            if t == REVOLUTE:
                if not (j == 4 or j == 10) or knee_rot:
                    pose[o] = pose[o] + draw()
            elif t == SPHERICAL:
                if j not in (3, 5, 9, 11):
                    rr = euler_to_quat([draw(), draw(), draw()])
                    pose[o:o + 4] = quat_mul(rr, pose[o:o + 4])
                    if rot_vel_w_pose:
                        vel[o:o + 4] = quat_mul(rr, vel[o:o + 4])
        if vel_noise:
            rr = euler_to_quat([draw(), draw(), draw()])
            vel[3:7] = quat_mul(rr, vel[3:7])
            for j in range(1, len(sizes)):
                o, t = offs[j], ch["joint_type"][j]
                if t == REVOLUTE:
                    if (not (j == 4 or j != 10)) or knee_rot:
                        vel[o] = vel[o] + draw()
                elif t == SPHERICAL:
                    if j not in (3, 5, 9, 11):
                        rr = euler_to_quat([draw(), draw(), draw()])
                        vel[o:o + 4] = quat_mul(rr, vel[o:o + 4])
        # cKinTree::PostProcessPose (KinTree.cpp:1558-1575)
        pose[3:7] = pose[3:7] / np.linalg.norm(pose[3:7])
        for j in range(1, len(sizes)):
            if ch["joint_type"][j] == SPHERICAL:
                o = offs[j]
                pose[o:o + 4] = pose[o:o + 4] / np.linalg.norm(pose[o:o + 4])

    if noise_bef_rot:
        add_noise_pose_vel()
        random_rotate()
    else:
        random_rotate()
        add_noise_pose_vel()
    return pose, vel, used[0]
