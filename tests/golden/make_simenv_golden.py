"""Writes tests/golden/simenv_golden.npz by running the REFERENCE's own gym_simenv/envs/sim_env.py.

Run in the build container only (needs /root/reference):  python tests/golden/make_simenv_golden.py

sim_env.py imports `gym` and the SWIG simulator (`deepmimic.env.deepmimic_env.DeepMimicEnv`), neither of which
exists in the image.  Only what those two provide is stubbed — gym's Env / spaces.Box / utils.EzPickle /
seeding.np_random base classes, and a stand-in simulator object that answers the size / rate getters and hands
back initial states — so that the reference's SimEnv.__init__ (with the reference's own
args/run_amp_humanoid3d_spinkick_args.txt, humanoid3d.txt and humanoid3d_rot_ctrl.txt), step, is_done,
check_collision / check_sphere / check_capsule / check_velocity and reset run UNMODIFIED.  The reference's
deepmimic/util/arg_parser.py and milo/milo/dynamics.py are imported as they are.  Initial states are the
env-state features of the reference clip (oracle.imitation_oracle.record_state, itself pinned to the reference's
kinematics code by tests/test_imitation_ref.py) — they are inputs, stored in the file.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import make_golden as mg  # noqa: E402  (reference milo loader + synthetic dataset)
from oracle import imitation_oracle as io  # noqa: E402
from tests import helpers as H  # noqa: E402

REF = mg.REF
DM_ROOT = os.path.join(REF, "deepmimic", "deepmimic")
ARG_FILE = "args/run_amp_humanoid3d_spinkick_args.txt"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simenv_golden.npz")
S, A = 226, 28


class FakeSimulator:
    """Stands in for DeepMimicEnv (deepmimic_env.py:6-313): getters + a queue of initial states."""
    clip = None
    log = []

    def __init__(self, args, enable_draw):
        self.args, self.time = args, None

    def set_mode(self, mode):
        pass

    def seed(self, s):
        FakeSimulator.log.append(("seed", s))

    def get_state_size(self, agent_id):
        return S

    def get_action_size(self, agent_id):
        return A

    def get_motion_length(self):
        return float(FakeSimulator.clip.duration)

    def get_agent_update_rate(self):
        return 30.0

    def get_pos_feature_dim(self):
        return 3

    def get_rot_feature_dim(self):
        return 6

    def get_vel_offset(self):
        return 136

    def reset_time(self, time, **kw):
        self.time = time

    def record_state(self, agent_id):
        c = FakeSimulator.clip
        return io.record_state(io.HUMANOID3D, c.kin_pose(self.time), c.kin_vel(self.time))


def load_reference_simenv():
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")

    class Env:
        pass

    class Box:
        def __init__(self, low, high, dtype):
            self.low, self.high, self.dtype, self.shape = low, high, dtype, low.shape

    class EzPickle:
        def __init__(self, *a, **kw):
            pass

    def np_random(seed=None):
        return np.random.RandomState(seed), seed

    gym.Env, gym.spaces, gym.utils = Env, spaces, utils
    spaces.Box = Box
    utils.EzPickle, utils.seeding = EzPickle, seeding
    seeding.np_random = np_random
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding})

    dm = types.ModuleType("deepmimic")
    dm.__path__ = []
    env = types.ModuleType("deepmimic.env")
    env.__path__ = []
    dme = types.ModuleType("deepmimic.env.deepmimic_env")
    dme.DeepMimicEnv = FakeSimulator
    util = types.ModuleType("deepmimic.util")
    util.__path__ = []
    spec = importlib.util.spec_from_file_location("deepmimic.util.arg_parser", os.path.join(DM_ROOT, "util", "arg_parser.py"))
    ap = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ap)   # the reference's own ArgParser
    sys.modules.update({"deepmimic": dm, "deepmimic.env": env, "deepmimic.env.deepmimic_env": dme,
                        "deepmimic.util": util, "deepmimic.util.arg_parser": ap})
    spec = importlib.util.spec_from_file_location("ref_sim_env", os.path.join(REF, "gym-simenv", "gym_simenv", "envs", "sim_env.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.SimEnv


def main():
    ref = mg.load_reference()
    SimEnv = load_reference_simenv()
    FakeSimulator.clip = io.Clip(H.spinkick_raw(), io.HUMANOID3D, "wrap")
    os.chdir(DM_ROOT)  # the arg file names its character / controller files relative to this directory

    N, hidden = 3, [32, 32]
    s, a, s2 = mg.synth_dataset(512, S, A, 0)
    ds = ref["datasets"].AmpDataset(s, a, s2)
    ens = mg.build_ensemble(ref, S, A, N, hidden, True, "relu", ds)
    out = {"N": np.int64(N), "hidden": np.array(hidden), "dataset_seed": np.int64(0), "dataset_rows": np.int64(512),
           "base_seed": np.int64(100),
           "weight_checksum": np.array([[float(l.weight.double().abs().sum()) for l in m.model.fc_layers]
                                        for m in ens.models])}

    # ---- A: episodes through reset / step / is_done with the horizon cut (horizon = 8)
    env = SimEnv(ens, deepmimic_args=ARG_FILE, horizon=8, seed=7)
    assert env.fall_contact_bodies_shapes.count("sphere") + env.fall_contact_bodies_shapes.count("capsule") == 13
    out.update(record_world_root_rot=np.int64(env.record_world_root_rot), record_all_world=np.int64(env.record_all_world),
               record_world_root_pos=np.int64(env.record_world_root_pos), time_max=np.float64(env.time_max),
               fall_offsets=np.asarray(env.fall_contact_bodies_offset),
               fall_params=np.asarray(env.fall_contact_bodies_params, dtype=np.float64),
               fall_is_capsule=np.array([sh == "capsule" for sh in env.fall_contact_bodies_shapes]))
    rng = np.random.default_rng(21)
    episodes, steps = 5, 10
    ob0 = np.zeros((episodes, S))
    acts = rng.normal(0, 1, (episodes, steps, A)).astype(np.float32).astype(np.float64)
    obs = np.zeros((episodes, steps, S))
    dones = np.zeros((episodes, steps), dtype=bool)
    nsteps = np.zeros((episodes, steps), dtype=np.int64)
    members = np.zeros(episodes, dtype=np.int64)
    assert env.dynamics is ens.models[0]  # sim_env.py:118-119
    for e in range(episodes):
        ob0[e] = env.reset()
        members[e] = env.reset_counter
        assert env.dynamics is ens.models[members[e]]
        for k in range(steps):   # keeps stepping past `done`, as a caller may
            ob, reward, done, info = env.step(acts[e, k].copy())
            assert reward == 0 and info == {}
            obs[e, k], dones[e, k], nsteps[e, k] = ob, done, env.num_steps
    out.update(traj_ob0=ob0, traj_actions=acts, traj_obs=obs, traj_done=dones, traj_num_steps=nsteps,
               traj_member=members)

    # ---- B: fall-contact thresholds, body by body, on both sides of `<= radius + 1e-4`
    base = ob0[0].copy()
    cases, flags, per_body = [], [], []
    for i in range(13):
        off = int(env.fall_contact_bodies_offset[i])
        radius = 0.5 * env.fall_contact_bodies_params[i][0]
        height = env.fall_contact_bodies_params[i][1]
        variants = [(0.0, +1)] if env.fall_contact_bodies_shapes[i] == "sphere" else [(0.8, +1), (0.8, -1), (-0.6, +1)]
        for norm_y, which in variants:
            for eps in (-1e-3, -2e-6, 2e-6, 1e-3):
                st = base.copy()
                target = radius + 1e-4 + eps          # world height of the lowest sphere / cap centre
                if env.fall_contact_bodies_shapes[i] == "sphere":
                    st[off + 1] = target - st[0]
                else:
                    st[off + 3 + 1] = norm_y
                    cap = 0.5 * height * norm_y * which
                    # put the chosen cap centre on the target; the other cap then sits higher when cap < 0
                    st[off + 1] = target - cap - st[0]
                env.set_observation(st)
                cases.append(st)
                flags.append(bool(env.check_collision()))
                per_body.append([bool(env.check_sphere(b)) if env.fall_contact_bodies_shapes[b] == "sphere"
                                 else bool(env.check_capsule(b)) for b in range(13)])
    out.update(contact_states=np.array(cases), contact_collided=np.array(flags), contact_per_body=np.array(per_body))

    # ---- C: velocity check (off by default, sim_env.py:27) and its threshold
    envv = SimEnv(ens, deepmimic_args=ARG_FILE, enable_velocity_check=True, horizon=300, seed=None)
    vstates, vflags, vdefault = [], [], []
    for idx, val in ((136, 100.0), (136, 100.0001), (225, -100.0001), (200, 99.9999), (135, 500.0), (150, -1e4)):
        st = base.copy()
        st[idx] = val
        envv.set_observation(st.copy())
        envv.num_steps = 1
        vflags.append(bool(envv.is_done()))
        env.set_observation(st.copy())
        env.num_steps = 1
        vdefault.append(bool(env.is_done()))
        vstates.append(st)
    out.update(vel_states=np.array(vstates), vel_done_enabled=np.array(vflags), vel_done_default=np.array(vdefault))

    # ---- D: the controller-file flags the reference honours (sim_env.py:92-96, 181-186, 224-231, 264-267), set on
    # the constructed object the way another controller file would set them
    rng = np.random.default_rng(33)
    flag_sets = [(1, 0), (0, 1), (1, 1)]
    var_states = np.zeros((len(flag_sets), 48, S))
    var_flags = np.zeros((len(flag_sets), 48), dtype=bool)
    for k, (aw, wrp) in enumerate(flag_sets):
        env.record_all_world, env.record_world_root_pos = bool(aw), bool(wrp)
        for i in range(48):
            st = base.copy()
            st[0] += rng.normal(0, 0.05)
            world_bodies = range(15) if aw else [0]     # bodies whose height this layout stores in world coordinates
            for bdy in world_bodies:
                st[9 * bdy + 2] += st[0]
            j = int(rng.integers(0, 13))                 # one fall body near its threshold, on either side
            body = int(env.fall_contact_bodies[j])
            off = int(env.fall_contact_bodies_offset[j])
            radius = 0.5 * env.fall_contact_bodies_params[j][0]
            eps = float(rng.choice([-1e-3, -2e-6, 2e-6, 1e-3]))
            cap = 0.0
            if env.fall_contact_bodies_shapes[j] == "capsule":
                cap = abs(0.5 * env.fall_contact_bodies_params[j][1] * st[off + 3 + 1])
            y_world = radius + 1e-4 + eps + cap          # the lower cap centre (or the sphere centre) on the target
            st[off + 1] = y_world if (aw or (wrp and j == 0)) else y_world - st[0]
            env.set_observation(st.copy())
            var_states[k, i] = st
            var_flags[k, i] = bool(env.check_collision())
    env.record_all_world, env.record_world_root_pos = False, False
    envv.record_vel_as_pos = True                        # velocities stored as per-step displacements: v = ob / dt
    vp_states, vp_flags = [], []
    for idx, val in ((136, 3.3), (136, 3.34), (225, -3.4), (200, 3.333), (150, 50.0)):
        st = base.copy()
        st[136:] *= envv.sampling_rate                   # the base state in that convention
        st[idx] = val
        envv.set_observation(st.copy())
        envv.num_steps = 1
        vp_flags.append(bool(envv.is_done()))
        vp_states.append(st)
    out.update(variant_states=var_states, variant_flag_sets=np.array(flag_sets), variant_collided=var_flags,
               velpos_states=np.array(vp_states), velpos_done=np.array(vp_flags),
               velpos_divisor=np.float64(envv.sampling_rate))

    np.savez_compressed(OUT, **out)
    print("variants collided:", var_flags.sum(1), "velpos:", vp_flags, envv.sampling_rate)
    print(OUT, os.path.getsize(OUT), "bytes; done per episode:", dones.argmax(1), "collided cases:", int(np.sum(flags)),
          "of", len(flags), "; velocity:", vflags, vdefault)


if __name__ == "__main__":
    main()
