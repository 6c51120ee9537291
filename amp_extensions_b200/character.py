"""Articulated-character tables for the imitation reward (DeepMimic character files).

A `Character` holds what reference DeepMimicCore reads from a character JSON file ("Skeleton"/"Joints" and
"BodyDefs", e.g. data/characters/humanoid3d.txt) and needs for cSceneImitate::CalcRewardImitate: joint
types, parents, attach points, pose/vel parameter offsets (anim/KinTree.cpp:824-850, 1053-1062),
end-effector flags, DiffWeight, body masses and body attach points.  `humanoid3d()` returns the built-in
table of the reference's humanoid; `from_json` parses any DeepMimic character file
(attach rotations must be zero, as they are for every shipped character used by this path).
"""
import json

import numpy as np

from . import _lib

JOINT_ROOT, JOINT_SPHERICAL, JOINT_REVOLUTE, JOINT_FIXED = 0, 1, 2, 3
_TYPE_IDS = {"none": JOINT_ROOT, "spherical": JOINT_SPHERICAL, "revolute": JOINT_REVOLUTE, "fixed": JOINT_FIXED}
_PARAM_SIZE = {JOINT_ROOT: 7, JOINT_SPHERICAL: 4, JOINT_REVOLUTE: 1, JOINT_FIXED: 0}


class Character:
    def __init__(self, names, joint_type, parent, attach, is_end_eff, diff_weight, body_mass, body_attach):
        self.names = list(names)
        self.joint_type = np.asarray(joint_type, dtype=np.int32)
        self.parent = np.asarray(parent, dtype=np.int32)
        self.attach = np.asarray(attach, dtype=np.float64).reshape(-1, 3).copy()
        self.attach[0] = 0.0  # the root's attach point is zeroed at load (KinTree.cpp:1064-1067)
        self.is_end_eff = np.asarray(is_end_eff, dtype=np.int32)
        self.diff_weight = np.asarray(diff_weight, dtype=np.float64)
        self.body_mass = np.asarray(body_mass, dtype=np.float64)
        self.body_attach = np.asarray(body_attach, dtype=np.float64).reshape(-1, 3)
        self.n_joints = len(self.names)
        if self.n_joints > _lib.MAX_JOINTS:
            raise ValueError(f"at most {_lib.MAX_JOINTS} joints are supported")
        if self.joint_type[0] != JOINT_ROOT or self.parent[0] != -1:
            raise ValueError("joint 0 must be the root")
        if any(self.parent[j] >= j for j in range(1, self.n_joints)):
            raise ValueError("joints must be topologically ordered (parent id < joint id)")
        sizes = np.array([_PARAM_SIZE[int(t)] for t in self.joint_type], dtype=np.int32)
        self.param_size = sizes
        self.param_offset = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
        self.dof = int(sizes.sum())
        self.body_rotation_ignored = False

    @classmethod
    def from_json(cls, path_or_dict, ignore_body_rotation=False):
        """ignore_body_rotation: accept bodies with attach rotations (e.g. the reference's dog3d neck and tail).  They
        turn the body's own frame, not its centre of mass (KinTree.cpp:1156-1166), so the imitation reward and the clip
        sampling are unaffected; the env-state features (record_state) would be, and refuse such a character."""
        d = path_or_dict
        if not isinstance(d, dict):
            with open(d) as f:
                d = json.load(f)
        joints = d["Skeleton"]["Joints"]
        bodies = {b["ID"]: b for b in d["BodyDefs"]}
        for j in joints:
            if any(abs(j.get(k, 0.0)) > 0 for k in ("AttachThetaX", "AttachThetaY", "AttachThetaZ")):
                raise NotImplementedError("joint attach rotations are not supported")
        rotated = any(abs(b.get(k, 0.0)) > 0 for b in bodies.values() for k in ("AttachThetaX", "AttachThetaY", "AttachThetaZ"))
        if rotated and not ignore_body_rotation:
            raise NotImplementedError("body attach rotations are not supported (pass ignore_body_rotation=True for the "
                                      "imitation reward, which does not depend on them)")
        ch = cls(
            names=[j["Name"] for j in joints],
            joint_type=[_TYPE_IDS[j["Type"]] for j in joints],
            parent=[j["Parent"] for j in joints],
            attach=[[j["AttachX"], j["AttachY"], j["AttachZ"]] for j in joints],
            is_end_eff=[int(j.get("IsEndEffector", 0)) for j in joints],
            diff_weight=[float(j.get("DiffWeight", 1.0)) for j in joints],
            body_mass=[float(bodies[j["ID"]]["Mass"]) if j["ID"] in bodies else 0.0 for j in joints],
            body_attach=[[bodies[j["ID"]]["AttachX"], bodies[j["ID"]]["AttachY"], bodies[j["ID"]]["AttachZ"]]
                         if j["ID"] in bodies else [0.0, 0.0, 0.0] for j in joints])
        ch.body_rotation_ignored = rotated
        return ch

    def joint_weights(self):
        """cSceneImitate::CalcJointWeights (SceneImitate.cpp:300-312): DiffWeight / sum |DiffWeight|."""
        return self.diff_weight / np.abs(self.diff_weight).sum()

    def to_struct(self):
        s = _lib.SimstepCharacter()
        s.n_joints = self.n_joints
        s.dof = self.dof
        for j in range(self.n_joints):
            s.joint_type[j] = int(self.joint_type[j])
            s.parent[j] = int(self.parent[j])
            s.param_offset[j] = int(self.param_offset[j])
            s.is_end_eff[j] = int(self.is_end_eff[j])
            s.diff_weight[j] = float(self.diff_weight[j])
            s.body_mass[j] = float(self.body_mass[j])
            for k in range(3):
                s.attach[j][k] = float(self.attach[j][k])
                s.body_attach[j][k] = float(self.body_attach[j][k])
        return s


def humanoid3d():
    """The reference's humanoid3d character."""
    S, R, F = JOINT_SPHERICAL, JOINT_REVOLUTE, JOINT_FIXED
    rows = [
        # name, type, parent, joint attach, EE, DiffW, mass, body attach
        ("root", JOINT_ROOT, -1, (0, 0, 0), 0, 1.0, 6.0, (0, 0.07, 0)),
        ("chest", S, 0, (0, 0.236151, 0), 0, 0.5, 14.0, (0, 0.12, 0)),
        ("neck", S, 1, (0, 0.223894, 0), 0, 0.3, 2.0, (0, 0.175, 0)),
        ("right_hip", S, 0, (0, 0, 0.084887), 0, 0.5, 4.5, (0, -0.21, 0)),
        ("right_knee", R, 3, (0, -0.421546, 0), 0, 0.3, 3.0, (0, -0.2, 0)),
        ("right_ankle", S, 4, (0, -0.40987, 0), 1, 0.2, 1.0, (0.045, -0.0225, 0)),
        ("right_shoulder", S, 1, (-0.02405, 0.2435, 0.18311), 0, 0.3, 1.5, (0, -0.14, 0)),
        ("right_elbow", R, 6, (0, -0.274788, 0), 0, 0.2, 1.0, (0, -0.12, 0)),
        ("right_wrist", F, 7, (0, -0.258947, 0), 1, 0.0, 0.5, (0, 0, 0)),
        ("left_hip", S, 0, (0, 0, -0.084887), 0, 0.5, 4.5, (0, -0.21, 0)),
        ("left_knee", R, 9, (0, -0.421546, 0), 0, 0.3, 3.0, (0, -0.2, 0)),
        ("left_ankle", S, 10, (0, -0.40987, 0), 1, 0.2, 1.0, (0.045, -0.0225, 0)),
        ("left_shoulder", S, 1, (-0.02405, 0.2435, -0.18311), 0, 0.3, 1.5, (0, -0.14, 0)),
        ("left_elbow", R, 12, (0, -0.274788, 0), 0, 0.2, 1.0, (0, -0.12, 0)),
        ("left_wrist", F, 13, (0, -0.258947, 0), 1, 0.0, 0.5, (0, 0, 0)),
    ]
    return Character(*[list(col) for col in zip(*rows)])
