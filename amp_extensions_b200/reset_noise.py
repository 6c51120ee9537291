"""Reset noise of the kinematic character, batched on the device.

`SimEnv.reset()` (gym-simenv/gym_simenv/envs/sim_env.py:270-285) resets through `reset_time(time, resolve, noise_bef_rot,
low, high, radian, rot_vel_w_pose, vel_noise, interp, knee_rot)`: the kinematic character is posed on the clip at `time`
and `cKinCharacter::AddNoise` (DeepMimicCore/anim/KinCharacter.cpp:340-532) perturbs its generalized pose and velocity
before the env state is recorded.  `add_reset_noise` is that perturbation for E characters at once, as tensor operations
on the device the poses already live on; the draws come from a torch generator (the reference draws from its C++ global
generator, which no batched code can replay) or are passed in explicitly (tests feed the oracle the same draws).

`resolve` (cSceneImitate::ResolveCharGroundIntersect) moves the SIMULATED character out of the ground with Bullet's
geometry and is not part of this path.
"""
import math

import torch

from .character import JOINT_REVOLUTE, JOINT_SPHERICAL

# the reference's hard-coded joint ids (KinCharacter.cpp:433, 447, 506, 518): knees get noise only with knee_rot, hips
# and ankles never
_KNEES = (4, 10)
_HIPS_ANKLES = (3, 5, 9, 11)


def _quat_mul(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], dim=-1)


def euler_to_quat(e):
    """cMathUtil::EulerToQuaternion (MathUtil.cpp:423-429) = EulerToAxisAngle (:347-378) + AxisAngleToQuaternion
    (:455-466); e [..., 3] -> quaternion (w, x, y, z) [..., 4]."""
    x, y, z = e.unbind(-1)
    xs, xc, ys, yc, zs, zc = x.sin(), x.cos(), y.sin(), y.cos(), z.sin(), z.cos()
    c = ((yc * zc + xs * ys * zs + xc * zc + xc * yc - 1) * 0.5).clamp(-1.0, 1.0)
    theta = torch.acos(c)
    m21 = xs * yc - xc * ys * zs + xs * zc
    m02 = xc * ys * zc + xs * zs + ys
    m10 = yc * zs - xs * ys * zc + xc * zs
    denom = torch.sqrt(m21 * m21 + m02 * m02 + m10 * m10)
    small = theta.abs() < 0.00001
    safe = torch.where(small, torch.ones_like(denom), denom)
    axis = torch.stack([torch.where(small, torch.zeros_like(m21), m21 / safe),
                        torch.where(small, torch.zeros_like(m02), m02 / safe),
                        torch.where(small, torch.ones_like(m10), m10 / safe)], dim=-1)
    h = 0.5 * theta
    return torch.cat([h.cos().unsqueeze(-1), h.sin().unsqueeze(-1) * axis], dim=-1)


def draw_layout(character, vel_noise=False, knee_rot=False):
    """Which of RandomRotatePoseVel's draws exist, in the reference's order: a list of (kind, joint) with kind in
    'yaw', 'pose1', 'pose3', 'vroot3', 'vel1', 'vel3' ('…3' entries consume three values)."""
    out = [("yaw", 0)]
    for j in range(1, character.n_joints):
        t = int(character.joint_type[j])
        if t == JOINT_REVOLUTE and (j not in _KNEES or knee_rot):
            out.append(("pose1", j))
        elif t == JOINT_SPHERICAL and j not in _HIPS_ANKLES:
            out.append(("pose3", j))
    if vel_noise:
        out.append(("vroot3", 0))
        for j in range(1, character.n_joints):
            t = int(character.joint_type[j])
            if t == JOINT_REVOLUTE and ((not (j == 4 or j != 10)) or knee_rot):   # sic, KinCharacter.cpp:506
                out.append(("vel1", j))
            elif t == JOINT_SPHERICAL and j not in _HIPS_ANKLES:
                out.append(("vel3", j))
    return out


def num_rotation_draws(character, vel_noise=False, knee_rot=False):
    return sum(3 if k.endswith("3") else 1 for k, _ in draw_layout(character, vel_noise, knee_rot))


def add_reset_noise(character, pose, vel, noise_bef_rot=False, noise_min=0.0, noise_max=0.0, radian=0.0,
                    rot_vel_w_pose=False, vel_noise=False, interp=1.0, knee_rot=False, generator=None, draws=None):
    """cKinCharacter::AddNoise for pose, vel [E, dof] (returns new tensors).  draws: optional (u_pose [E, dof],
    u_vel [E, dof] in [0, 1); r [E, num_rotation_draws] in [-1, 1)) instead of the generator."""
    E, dof = pose.shape
    dev, dt = pose.device, pose.dtype
    n_r = num_rotation_draws(character, vel_noise, knee_rot)
    if draws is None:
        u_pose = torch.rand((E, dof), device=dev, dtype=dt, generator=generator)
        u_vel = torch.rand((E, dof), device=dev, dtype=dt, generator=generator)
        r = torch.rand((E, n_r), device=dev, dtype=dt, generator=generator) * 2 - 1
    else:
        u_pose, u_vel, r = (torch.as_tensor(t).to(dev, dt) for t in draws)
    pose, vel = pose.clone(), vel.clone()
    off = [int(o) for o in character.param_offset]

    def add_noise_pose_vel():                                     # KinCharacter.cpp:354-366
        if noise_min == 0 and noise_max == 0:
            return
        pose.add_(noise_min + (noise_max - noise_min) * u_pose)
        vel.add_(noise_min + (noise_max - noise_min) * u_vel)

    def random_rotate():                                          # KinCharacter.cpp:367-532
        if radian == 0:
            return
        k = 0
        for kind, j in draw_layout(character, vel_noise, knee_rot):
            o = off[j]
            if kind == "yaw":
                yaw = radian * r[:, k]
                k += 1
                rot = torch.stack([(0.5 * yaw).cos(), torch.zeros_like(yaw), (0.5 * yaw).sin(), torch.zeros_like(yaw)], -1)
                q = _quat_mul(rot, pose[:, 3:7])                  # cCharacter::RotateRoot (Character.cpp:210-216)
                pose[:, 3:7] = q / q.norm(dim=-1, keepdim=True)
                vel.mul_(interp)                                  # root, root angular and every joint's velocity (:398-417)
            elif kind == "pose1":
                pose[:, o] += radian * r[:, k]
                k += 1
            elif kind == "pose3":
                rr = euler_to_quat(radian * r[:, k:k + 3])
                k += 3
                pose[:, o:o + 4] = _quat_mul(rr, pose[:, o:o + 4])
                if rot_vel_w_pose:
                    vel[:, o:o + 4] = _quat_mul(rr, vel[:, o:o + 4])
            elif kind == "vroot3":
                rr = euler_to_quat(radian * r[:, k:k + 3])
                k += 3
                vel[:, 3:7] = _quat_mul(rr, vel[:, 3:7])          # the 4 root angular-velocity slots as (w, x, y, z) (:483-489)
            elif kind == "vel1":
                vel[:, o] += radian * r[:, k]
                k += 1
            elif kind == "vel3":
                rr = euler_to_quat(radian * r[:, k:k + 3])
                k += 3
                vel[:, o:o + 4] = _quat_mul(rr, vel[:, o:o + 4])
        # cKinTree::PostProcessPose (KinTree.cpp:1558-1575): unit root and spherical-joint quaternions
        pose[:, 3:7] = pose[:, 3:7] / pose[:, 3:7].norm(dim=-1, keepdim=True)
        for j in range(1, character.n_joints):
            if int(character.joint_type[j]) == JOINT_SPHERICAL:
                o = off[j]
                pose[:, o:o + 4] = pose[:, o:o + 4] / pose[:, o:o + 4].norm(dim=-1, keepdim=True)

    if noise_bef_rot:
        add_noise_pose_vel()
        random_rotate()
    else:
        random_rotate()
        add_noise_pose_vel()
    return pose, vel


def reset_kwargs(reset_args):
    """The keyword arguments of add_reset_noise from a SimEnv `reset_args` dict (sim_env.py:29-31, 78-81)."""
    if not reset_args:
        return None
    return dict(noise_bef_rot=bool(reset_args.get("noise_bef_rot", False)),
                noise_min=float(reset_args.get("noise_min", 0.0)), noise_max=float(reset_args.get("noise_max", 0.0)),
                radian=float(reset_args.get("radian", 0.0)), rot_vel_w_pose=bool(reset_args.get("rot_vel_w_pose", False)),
                vel_noise=bool(reset_args.get("vel_noise", False)), interp=float(reset_args.get("interp", 1.0)),
                knee_rot=bool(reset_args.get("knee_rot", False)))
