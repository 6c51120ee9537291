"""Writes tests/golden/imitation_ref_golden.npz: outputs of the reference's OWN DeepMimicCore kinematics code.

Run in the build container (needs /root/reference):  python tests/golden/make_imitation_ref_golden.py

oracle/ref_build.py compiles the reference's util/MathUtil.cpp, anim/KinTree.cpp, anim/Motion.cpp, sim/RBDUtil.cpp,
sim/SpAlg.cpp, ... where they lie, against the Eigen stand-in (oracle/eigen_shim; Eigen 3.3.7 itself is absent), into
oracle/_ref/libdmref.so; this script feeds it seeded float32-representable inputs and stores inputs + outputs.  The
vectors travel to the GPU box (the reference tree does not) and pin both the float64 restatement
(tests/test_imitation_ref.py) and the CUDA kernels (tests/test_imitation_gpu.py).
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from oracle import imitation_oracle as io  # noqa: E402  (input generation only: perturbs clip poses)
from oracle import ref_build as rb  # noqa: E402
from tests import helpers as H  # noqa: E402

PD = ctypes.POINTER(ctypes.c_double)


def P(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(PD)


def f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32).astype(np.float64))


def main():
    lib = rb.load()
    assert lib is not None and rb.reference_present(), "needs the reference tree"
    clip = io.Clip(H.spinkick_raw(), io.HUMANOID3D, "wrap")
    dof, nj, nf = lib.dmref_num_dof(), lib.dmref_num_joints(), lib.dmref_num_frames()
    out = {}

    # ---- the clip as the reference loads it
    frames, vels, times = np.zeros((nf, dof)), np.zeros((nf, dof)), np.zeros(nf)
    lib.dmref_clip_table(P(frames), P(vels), P(times))
    w = np.zeros(nj)
    lib.dmref_joint_weights(P(w))
    out.update(clip_frames=frames, clip_vels=vels, clip_times=times, clip_duration=np.float64(lib.dmref_duration()),
               clip_loop=np.int64(lib.dmref_loop()), joint_weights=w,
               param_offset=np.array([lib.dmref_param_offset(j) for j in range(nj)]),
               param_size=np.array([lib.dmref_param_size(j) for j in range(nj)]))

    # ---- clip sampling (cMotion::CalcFrame / CalcFrameVel + the kinematic character's loop offset and origin)
    rng = np.random.default_rng(11)
    ts = f32(np.concatenate([rng.uniform(-1.0, 6.0, 80), times[:8], [0.0, 1.283282, 2.566564]]))
    org = f32(rng.normal(0, 0.4, (ts.size, 3)))
    kp, kv = np.zeros((ts.size, dof)), np.zeros((ts.size, dof))
    for e in range(ts.size):
        lib.dmref_kin_pose_vel(float(ts[e]), P(org[e].copy()), P(kp[e]), P(kv[e]))
    out.update(sample_t=ts, sample_origin=org, sample_pose=kp, sample_vel=kv)

    # ---- reward: small and large perturbations, with and without a kinematic-character origin
    for tag, with_origin, seed, scale in (("plain", False, 3, 1.0), ("origin", True, 4, 1.0), ("far", False, 5, 4.0)):
        E = 96
        pose, vel, t, origin = H.perturbed_poses(E, seed=seed, clip=clip, t_max=4 * clip.duration, with_origin=with_origin)
        if scale != 1.0:
            r2 = np.random.default_rng(seed + 100)
            pose[:, 0:3] += r2.normal(0, 0.2, (E, 3))
            vel += r2.normal(0, 2.0, vel.shape)
            offs, sizes = io.param_layout(io.HUMANOID3D)
            for j in range(nj):
                if sizes[j] == 4:
                    vel[:, offs[j] + 3] = 0.0
            vel[:, 6] = 0.0
        pose, vel, t = f32(pose), f32(vel), f32(t)
        origin = f32(origin) if with_origin else None
        r, terms = np.zeros(E), np.zeros((E, 5))
        lib.dmref_reward_batch(E, P(pose), P(vel), P(t), P(origin) if with_origin else None, P(r), P(terms))
        out.update({f"{tag}_pose": pose, f"{tag}_vel": vel, f"{tag}_t": t, f"{tag}_reward": r, f"{tag}_terms": terms})
        if with_origin:
            out[f"{tag}_origin"] = origin

    # ---- building blocks on the "plain" inputs: per-joint errors, FK, heading / origin transform, COM
    pose, vel, t = out["plain_pose"], out["plain_vel"], out["plain_t"]
    E = 24
    pose_err, vel_err = np.zeros((E, nj)), np.zeros((E, nj))
    jpos, jtrans = np.zeros((E, nj, 3)), np.zeros((E, nj, 16))
    heading, otrans = np.zeros(E), np.zeros((E, 16))
    com, com_vel = np.zeros((E, 3)), np.zeros((E, 3))
    for e in range(E):
        p1, v1 = np.zeros(dof), np.zeros(dof)
        lib.dmref_kin_pose_vel(float(t[e]), None, P(p1), P(v1))
        p0, v0 = pose[e].copy(), vel[e].copy()
        for j in range(nj):
            pose_err[e, j] = lib.dmref_pose_err(j, P(p0), P(p1))
            vel_err[e, j] = lib.dmref_vel_err(j, P(v0), P(v1))
            lib.dmref_joint_world_pos(P(p0), j, P(jpos[e, j]))
            lib.dmref_joint_world_trans(P(p0), j, P(jtrans[e, j]))
        heading[e] = lib.dmref_heading(P(p0))
        lib.dmref_origin_trans(P(p0), P(otrans[e]))
        lib.dmref_com(P(p0), P(v0), P(com[e]), P(com_vel[e]))
    out.update(blk_pose_err=pose_err, blk_vel_err=vel_err, blk_joint_pos=jpos, blk_joint_trans=jtrans,
               blk_heading=heading, blk_origin_trans=otrans, blk_com=com, blk_com_vel=com_vel)

    # ---- env state features (CtController::BuildStatePose / BuildStateVel over the reference's kinematic helpers)
    flag_sets = [(0, 0, 1, 1.0), (0, 0, 0, 1.0), (1, 0, 1, 1.0), (0, 1, 1, 1.0 / 30.0)]
    states = np.zeros((len(flag_sets), E, 1 + 9 * nj + 6 * nj))
    for k, (aw, wrp, wrr, vs) in enumerate(flag_sets):
        for e in range(E):
            lib.dmref_record_state(P(pose[e].copy()), P(vel[e].copy()), aw, wrp, wrr, vs, P(states[k, e]))
    out.update(state_flags=np.array(flag_sets), state_features=states)

    # ---- two more reference clips: a non-looping one (kick: times clamp, velocity 0 past the end) and a short cycle
    import json
    for name in ("humanoid3d_kick", "humanoid3d_run"):
        mfile = os.path.join(rb.DATA, "motions", name + ".txt")
        with open(mfile) as f:
            mj = json.load(f)
        raw = np.array(mj["Frames"], dtype=np.float64)
        loop = mj.get("Loop", "none")
        assert lib.dmref_init(rb.CHAR_FILE.encode(), mfile.encode()) == 0
        c2 = io.Clip(raw, io.HUMANOID3D, loop)
        nf2 = lib.dmref_num_frames()
        fr2, ve2, ti2 = np.zeros((nf2, dof)), np.zeros((nf2, dof)), np.zeros(nf2)
        lib.dmref_clip_table(P(fr2), P(ve2), P(ti2))
        dur = lib.dmref_duration()
        r3 = np.random.default_rng(len(name))
        ts2 = f32(np.concatenate([r3.uniform(-0.5, 2.5 * dur, 40), [0.0, dur, 0.5 * dur, dur + 0.25]]))
        org2 = f32(r3.normal(0, 0.3, (ts2.size, 3)))
        kp2, kv2 = np.zeros((ts2.size, dof)), np.zeros((ts2.size, dof))
        for e in range(ts2.size):
            lib.dmref_kin_pose_vel(float(ts2[e]), P(org2[e].copy()), P(kp2[e]), P(kv2[e]))
        E2 = 48
        pose2, vel2, t2, _ = H.perturbed_poses(E2, seed=len(name), clip=c2, t_max=2.0 * dur, with_origin=False)
        pose2, vel2, t2 = f32(pose2), f32(vel2), f32(t2)
        r2, terms2 = np.zeros(E2), np.zeros((E2, 5))
        lib.dmref_reward_batch(E2, P(pose2), P(vel2), P(t2), None, P(r2), P(terms2))
        out.update({f"{name}/raw": raw, f"{name}/loop": np.array(loop), f"{name}/frames": fr2, f"{name}/vels": ve2,
                    f"{name}/times": ti2, f"{name}/duration": np.float64(dur), f"{name}/sample_t": ts2,
                    f"{name}/sample_origin": org2, f"{name}/sample_pose": kp2, f"{name}/sample_vel": kv2,
                    f"{name}/pose": pose2, f"{name}/vel": vel2, f"{name}/t": t2, f"{name}/reward": r2,
                    f"{name}/terms": terms2})
    # ---- a different skeleton through the same code: the reference's dog (23 joints, 83 dof, four end effectors)
    cfile = os.path.join(rb.DATA, "characters", "dog3d.txt")
    mfile = os.path.join(rb.DATA, "motions", "dog3d_trot.txt")
    with open(cfile) as f:
        cj = json.load(f)
    with open(mfile) as f:
        mj = json.load(f)
    assert lib.dmref_init(cfile.encode(), mfile.encode()) == 0
    dog = H.character_dict_from_json(cj)
    ddof, dnj, dnf = lib.dmref_num_dof(), lib.dmref_num_joints(), lib.dmref_num_frames()
    raw = np.array(mj["Frames"], dtype=np.float64)
    loop = mj.get("Loop", "none")
    cd = io.Clip(raw, dog, loop)
    frd, ved, tid = np.zeros((dnf, ddof)), np.zeros((dnf, ddof)), np.zeros(dnf)
    lib.dmref_clip_table(P(frd), P(ved), P(tid))
    wd = np.zeros(dnj)
    lib.dmref_joint_weights(P(wd))
    dur = lib.dmref_duration()
    r4 = np.random.default_rng(44)
    tsd = f32(r4.uniform(-0.3, 3.0 * dur, 40))
    orgd = f32(r4.normal(0, 0.3, (tsd.size, 3)))
    kpd, kvd = np.zeros((tsd.size, ddof)), np.zeros((tsd.size, ddof))
    for e in range(tsd.size):
        lib.dmref_kin_pose_vel(float(tsd[e]), P(orgd[e].copy()), P(kpd[e]), P(kvd[e]))
    Ed = 48
    posed, veld, td, _ = H.perturbed_poses(Ed, seed=45, clip=cd, t_max=2.0 * dur, with_origin=False, ch=dog)
    posed, veld, td = f32(posed), f32(veld), f32(td)
    rd, termsd = np.zeros(Ed), np.zeros((Ed, 5))
    lib.dmref_reward_batch(Ed, P(posed), P(veld), P(td), None, P(rd), P(termsd))
    comd, comvd = np.zeros((8, 3)), np.zeros((8, 3))
    for e in range(8):
        lib.dmref_com(P(posed[e].copy()), P(veld[e].copy()), P(comd[e]), P(comvd[e]))
    out.update({"dog/character_json": np.array(json.dumps({"Skeleton": cj["Skeleton"], "BodyDefs": cj["BodyDefs"]})),
                "dog/raw": raw, "dog/loop": np.array(loop), "dog/frames": frd, "dog/vels": ved, "dog/times": tid,
                "dog/duration": np.float64(dur), "dog/joint_weights": wd, "dog/sample_t": tsd, "dog/sample_origin": orgd,
                "dog/sample_pose": kpd, "dog/sample_vel": kvd, "dog/pose": posed, "dog/vel": veld, "dog/t": td,
                "dog/reward": rd, "dog/terms": termsd, "dog/com": comd, "dog/com_vel": comvd})
    print("dog:", ddof, dnj, dnf, dur, loop, "reward range", rd.min(), rd.max())

    assert lib.dmref_init(rb.CHAR_FILE.encode(), rb.MOTION_FILE.encode()) == 0   # back to the spin kick

    path = os.path.join(ROOT, "tests", "golden", "imitation_ref_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", "reward range", out["plain_reward"].min(), out["far_reward"].min(),
          out["plain_reward"].max())


if __name__ == "__main__":
    main()
