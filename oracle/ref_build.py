"""TEST INFRASTRUCTURE ONLY: builds oracle/_ref/libdmref.so from the reference's own DeepMimicCore sources.

The recipe compiles the reference .cpp files WHERE THEY LIE under /root/reference (nothing is copied into this
repository) together with oracle/ref_driver.cpp, against the Eigen stand-in in oracle/eigen_shim (Eigen 3.3.7,
Bullet 2.88 and SWIG are absent from the image; only the kinematics files, which need Eigen alone, are built).
Outputs go to oracle/_ref/ only (git-ignored, shipped to the GPU box by gpurun): the library and, beside it under
oracle/_ref/data/, the two reference DATA files it parses at start-up (humanoid3d.txt, humanoid3d_spinkick.txt), so
that a box without /root/reference can still run the prebuilt library.  `python -m oracle.ref_build` or
`__graft_entry__.build()` run the recipe; without /root/reference it is a no-op.
"""
import ctypes
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("AMP_REFERENCE_ROOT", "/root/reference")
CORE = os.path.join(REF_ROOT, "deepmimic", "deepmimic", "DeepMimicCore")
DATA = os.path.join(REF_ROOT, "deepmimic", "deepmimic", "data")
CHAR_FILE = os.path.join(DATA, "characters", "humanoid3d.txt")
MOTION_FILE = os.path.join(DATA, "motions", "humanoid3d_spinkick.txt")
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libdmref.so")
# the stand-in headers, unless a real Eigen 3.3.7 tree is pointed to (none exists in the build image)
EIGEN_INCLUDE = os.environ.get("EIGEN3_INCLUDE_DIR") or os.path.join(HERE, "eigen_shim")
SHIPPED_CHAR = os.path.join(OUT_DIR, "data", "humanoid3d.txt")
SHIPPED_MOTION = os.path.join(OUT_DIR, "data", "humanoid3d_spinkick.txt")

# the reference translation units that make up the kinematic half of DeepMimicCore (no Bullet, no OpenGL)
REF_SOURCES = [
    "util/MathUtil.cpp", "util/Rand.cpp", "util/JsonUtil.cpp", "util/FileUtil.cpp",
    "util/json/json_reader.cpp", "util/json/json_value.cpp", "util/json/json_writer.cpp",
    "anim/Shape.cpp", "anim/KinTree.cpp", "anim/Motion.cpp",
    "sim/SpAlg.cpp", "sim/RBDModel.cpp", "sim/RBDUtil.cpp",
]


def reference_present():
    return os.path.isdir(CORE) and os.path.exists(CHAR_FILE) and os.path.exists(MOTION_FILE)


def build(force=False, verbose=False):
    """Returns the path of libdmref.so, or None when the reference tree is not on this machine."""
    if not reference_present():
        return LIB if os.path.exists(LIB) else None
    srcs = [os.path.join(CORE, s) for s in REF_SOURCES] + [os.path.join(HERE, "ref_driver.cpp")]
    deps = srcs + [os.path.join(HERE, "eigen_shim", "Eigen", "Core"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(OUT_DIR, os.path.relpath(s, CORE if s.startswith(CORE) else HERE).replace(os.sep, "_")[:-4] + ".o")
        objs.append(o)
        cmd = ["g++", "-std=c++14", "-O2", "-fPIC", "-w", "-ffp-contract=off", "-I", EIGEN_INCLUDE, "-I", CORE, "-c", s,
               "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stderr=subprocess.PIPE, text=True)))
    for s, p in procs:
        _, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"compiling {s} failed:\n{err[-4000:]}")
    subprocess.run(["g++", "-shared", "-o", LIB] + objs, check=True)
    for o in objs:
        os.remove(o)
    os.makedirs(os.path.dirname(SHIPPED_CHAR), exist_ok=True)
    shutil.copyfile(CHAR_FILE, SHIPPED_CHAR)
    shutil.copyfile(MOTION_FILE, SHIPPED_MOTION)
    return LIB


_lib = None


def load():
    """ctypes handle on libdmref.so, initialised with the reference's humanoid3d character and spinkick clip.
    Returns None when neither the library nor the reference data are available (e.g. on the GPU box)."""
    global _lib
    if _lib is not None:
        return _lib
    path = build()
    if path is None:
        return None
    if reference_present():
        char_file, motion_file = CHAR_FILE, MOTION_FILE
    elif os.path.exists(SHIPPED_CHAR) and os.path.exists(SHIPPED_MOTION):
        char_file, motion_file = SHIPPED_CHAR, SHIPPED_MOTION
    else:
        return None
    lib = ctypes.CDLL(path)
    d, i, pd = ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
    lib.dmref_init.argtypes, lib.dmref_init.restype = [ctypes.c_char_p, ctypes.c_char_p], i
    for name in ("dmref_num_dof", "dmref_num_joints", "dmref_num_frames", "dmref_loop"):
        getattr(lib, name).argtypes, getattr(lib, name).restype = [], i
    lib.dmref_duration.argtypes, lib.dmref_duration.restype = [], d
    lib.dmref_param_offset.argtypes, lib.dmref_param_offset.restype = [i], i
    lib.dmref_param_size.argtypes, lib.dmref_param_size.restype = [i], i
    lib.dmref_joint_weights.argtypes, lib.dmref_joint_weights.restype = [pd], None
    lib.dmref_clip_table.argtypes, lib.dmref_clip_table.restype = [pd, pd, pd], None
    lib.dmref_kin_pose_vel.argtypes, lib.dmref_kin_pose_vel.restype = [d, pd, pd, pd], None
    lib.dmref_pose_err.argtypes, lib.dmref_pose_err.restype = [i, pd, pd], d
    lib.dmref_vel_err.argtypes, lib.dmref_vel_err.restype = [i, pd, pd], d
    lib.dmref_joint_world_pos.argtypes, lib.dmref_joint_world_pos.restype = [pd, i, pd], None
    lib.dmref_joint_world_trans.argtypes, lib.dmref_joint_world_trans.restype = [pd, i, pd], None
    lib.dmref_heading.argtypes, lib.dmref_heading.restype = [pd], d
    lib.dmref_origin_trans.argtypes, lib.dmref_origin_trans.restype = [pd, pd], None
    lib.dmref_com.argtypes, lib.dmref_com.restype = [pd, pd, pd, pd], None
    lib.dmref_lerp_poses.argtypes, lib.dmref_lerp_poses.restype = [pd, pd, d, pd], None
    lib.dmref_calc_vel.argtypes, lib.dmref_calc_vel.restype = [pd, pd, d, pd], None
    lib.dmref_quat_theta.argtypes, lib.dmref_quat_theta.restype = [pd], d
    lib.dmref_quat_rot_vec.argtypes, lib.dmref_quat_rot_vec.restype = [pd, pd, pd], None
    lib.dmref_normal_tangent.argtypes, lib.dmref_normal_tangent.restype = [pd, pd, pd], None
    lib.dmref_reward.argtypes, lib.dmref_reward.restype = [pd, pd, pd, pd, d, pd], d
    lib.dmref_record_state.argtypes, lib.dmref_record_state.restype = [pd, pd, i, i, i, d, pd], None
    lib.dmref_reward_batch.argtypes, lib.dmref_reward_batch.restype = [i, pd, pd, pd, pd, pd, pd], None
    lib.dmref_reset_noise.argtypes = [pd, pd, i, d, d, d, i, i, d, i, pd, pd, pd, pd, pd]
    lib.dmref_reset_noise.restype = i
    # cMotion::Load reports on stdout (Motion.cpp); keep the caller's stdout clean (bench lines are parsed as JSON)
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        rc = lib.dmref_init(char_file.encode(), motion_file.encode())
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    if rc != 0:
        raise RuntimeError(f"dmref_init failed with code {rc}")
    _lib = lib
    return lib


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p if p else "reference tree not present: nothing built")
