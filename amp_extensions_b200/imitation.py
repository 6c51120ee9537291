"""DeepMimic / AMP motion-imitation reward against a reference clip, batched on B200.

`ImitationReward(character, clip)` evaluates reference DeepMimicCore
scenes/SceneImitate.cpp:7-127 (cSceneImitate::CalcRewardImitate) for E environments per call: the
simulated character is given by its generalized pose/velocity (the layout of cKinTree, e.g. 43 numbers
for humanoid3d: root pos 3 + root quat 4 (w,x,y,z) + per-joint parameters), the kinematic character is the
clip evaluated at each env's own time (Motion.cpp:267-305 with slerp, MotionController.cpp:25-41 cycle
offset, KinCharacter.cpp:573-640) plus an optional world offset of its origin.
"""
import torch

from . import engine as _engine
from . import reset_noise as _reset_noise
from .character import Character, humanoid3d
from .motion import MotionClip


class ImitationReward:
    def __init__(self, character=None, clip=None, device=None, engine=None):
        self.character = character if character is not None else humanoid3d()
        self.clip = clip if clip is not None else MotionClip.spinkick(self.character)
        if engine is None:
            # a handle needs an ensemble shape; the reward itself uses none of it
            engine = _engine.Engine(state_dim=self.character.dof, action_dim=0, num_models=1, hidden_sizes=[],
                                    transform=False, device=device)
        self.engine = engine
        self.engine.load_clip(self.character, self.clip)
        self.dof = self.character.dof

    def reward(self, pose, vel, kin_time, kin_origin=None, want_terms=False, out=None):
        """pose, vel: [E, dof]; kin_time: [E] seconds; kin_origin: [E, 3] or None.
        Returns reward [E] (written into `out` when given; and the five sub-rewards pose, vel, end-effector, root,
        com as [E, 5])."""
        return self.engine.imitation_reward(pose, vel, kin_time, kin_origin, want_terms=want_terms, out=out)

    def sample(self, kin_time, kin_origin=None):
        """Pose and velocity of the kinematic character at the given clip times: ([E, dof], [E, dof])."""
        return self.engine.clip_sample(kin_time, kin_origin)

    def record_state(self, pose, vel, **flags):
        """The env's state vector for generalized poses / velocities (CtController.cpp:378-495), e.g. 226 numbers
        for humanoid3d; flags as in Engine.record_state."""
        if getattr(self.character, "body_rotation_ignored", False):
            raise NotImplementedError("state features need the body attach rotations this character was loaded without")
        return self.engine.record_state(pose, vel, **flags)

    def reset_states(self, kin_time, kin_origin=None, reset_args=None, generator=None, draws=None, **flags):
        """Initial env states for clip times (sim_env.py:270-285): the character is set to the kinematic pose and
        velocity at `kin_time`, perturbed as `reset_args` says (the plugin's dict, sim_env.py:28-31: noise_bef_rot,
        noise_min / noise_max, radian, rot_vel_w_pose, vel_noise, interp, knee_rot -> cKinCharacter::AddNoise,
        KinCharacter.cpp:340-532, batched in reset_noise.add_reset_noise; `resolve` needs the simulator and is not
        applied) and its state is recorded.  generator: torch generator of the noise draws; draws: explicit draws."""
        pose, vel = self.sample(kin_time, kin_origin)
        kw = _reset_noise.reset_kwargs(reset_args)
        if kw is not None:
            pose, vel = _reset_noise.add_reset_noise(self.character, pose, vel, generator=generator, draws=draws, **kw)
        return self.record_state(pose, vel, **flags)

    @staticmethod
    def advance_time(kin_time, num_steps, dt=1.0 / 30.0):
        """Clip time after num_steps policy steps (20 substeps of 1/600 s, gym_deepmimic.py:106-110)."""
        return kin_time + torch.as_tensor(num_steps, dtype=torch.float32, device=kin_time.device) * dt
