// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI probe around the reference's OWN DeepMimicCore kinematics code.
//
// oracle/ref_build.py compiles this file together with the reference sources where they lie under
// /root/reference/deepmimic/deepmimic/DeepMimicCore (util/MathUtil.cpp, anim/KinTree.cpp, anim/Motion.cpp,
// sim/RBDUtil.cpp, sim/SpAlg.cpp, sim/RBDModel.cpp, util/JsonUtil.cpp, util/FileUtil.cpp, util/json/*.cpp, ...)
// into oracle/_ref/libdmref.so.  Eigen 3.3.7 is absent from the image; the sources are compiled against the
// stand-in headers in oracle/eigen_shim (see Eigen/Core there).  Everything numerically substantive below is
// a call INTO reference code: cKinTree::{CalcPoseErr, CalcVelErr, CalcRootRotErr, CalcRootAngVelErr,
// CalcJointWorldPos, JointWorldTrans, BuildOriginTrans, CalcHeading, LerpPoses, CalcVel, PostProcessPose},
// cRBDUtil::CalcCoM, cMotion::{Load, CalcFrame, CalcFrameVel}, cMathUtil::*.
//
// What cannot be compiled (needs Bullet / OpenGL) and is therefore restated here, each with its source lines:
//   * cSceneImitate::CalcRewardImitate  (scenes/SceneImitate.cpp:7-127)  - the ~60 lines that combine the error terms;
//     the simulated character's pose/COM queries go through the same cKinTree / cRBDUtil calls the reference uses
//     for the kinematic character (deviation documented in DESIGN.md section 4), ground height 0
//   * cSceneImitate::CalcJointWeights   (scenes/SceneImitate.cpp:300-312)
//   * cKinController::{BuildMotionParams, PostProcessMotion, CalcCycleRootDelta} (anim/KinController.cpp:101-175)
//   * cMotionController::CalcPose / CalcRootCycleOffset (anim/MotionController.cpp:25-47, :144-153)
//   * cKinCharacter::CalcPose / CalcVel (anim/KinCharacter.cpp:573-640) with mOriginRot = identity
//   * cCtController::BuildStatePose / BuildStateVel (sim/CtController.cpp:378-495): the loop that lays out the env
//     state; body positions / rotations / velocities, which the reference reads from Bullet bodies, come from the
//     reference's kinematic equivalents cKinTree::{BodyWorldTrans, CalcBodyPartVel, CalcJointWorldAngularVel}
//   * cKinCharacter::AddNoise / AddNoisePoseVel / RandomRotatePoseVel (anim/KinCharacter.cpp:340-532): the reset noise,
//     with the reference's random draws turned into inputs (dmref_reset_noise)
// Used by tests/test_imitation_ref.py and tests/golden/make_imitation_ref_golden.py; never by the product.
#include <cstring>
#include <fstream>
#include <string>

#include "anim/KinTree.h"
#include "anim/Motion.h"
#include "sim/RBDUtil.h"
#include "util/MathUtil.h"

namespace {

struct tState {
    bool ok = false;
    Eigen::MatrixXd joint_mat;
    Eigen::MatrixXd body_defs;
    Eigen::VectorXd joint_weights;
    cMotion motion;
    tVector cycle_root_delta;
    int dof = 0;
    int num_joints = 0;
};
tState g;

Eigen::VectorXd ToVec(const double* p, int n) {
    Eigen::VectorXd v(n);
    for (int i = 0; i < n; ++i) v[i] = p[i];
    return v;
}
void FromVec(const Eigen::VectorXd& v, double* out) {
    for (int i = 0; i < static_cast<int>(v.size()); ++i) out[i] = v[i];
}
void FromMat4(const tMatrix& m, double* out) {  // row-major 4x4
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[4 * i + j] = m(i, j);
}

// cKinCharacter::CalcPose (KinCharacter.cpp:573-598) over cMotionController::CalcPose (MotionController.cpp:25-41)
void KinPose(double time, const tVector& origin, Eigen::VectorXd& out_pose) {
    g.motion.CalcFrame(time, out_pose);
    if (g.motion.EnableLoop()) {
        int cycle_count = g.motion.CalcCycleCount(time);   // MotionController.cpp:144-153
        tVector root_pos = cKinTree::GetRootPos(out_pose);
        root_pos += cycle_count * g.cycle_root_delta;
        cKinTree::SetRootPos(root_pos, out_pose);
    }
    const tQuaternion origin_rot = tQuaternion::Identity();
    tVector root_pos = cKinTree::GetRootPos(out_pose);
    tQuaternion root_rot = cKinTree::GetRootRot(out_pose);
    root_rot = origin_rot * root_rot;
    root_rot = cMathUtil::StandardizeQuat(root_rot);
    root_pos = cMathUtil::QuatRotVec(origin_rot, root_pos);
    root_pos += origin;
    cKinTree::SetRootPos(root_pos, out_pose);
    cKinTree::SetRootRot(root_rot, out_pose);
}

// cKinCharacter::CalcVel (KinCharacter.cpp:622-640) over cMotionController::CalcVel (MotionController.cpp:43-47)
void KinVel(double time, Eigen::VectorXd& out_vel) {
    g.motion.CalcFrameVel(time, out_vel);
    const tQuaternion origin_rot = tQuaternion::Identity();
    tVector root_vel = cKinTree::GetRootVel(out_vel);
    tVector root_ang_vel = cKinTree::GetRootAngVel(out_vel);
    root_vel = cMathUtil::QuatRotVec(origin_rot, root_vel);
    root_ang_vel = cMathUtil::QuatRotVec(origin_rot, root_ang_vel);
    cKinTree::SetRootVel(root_vel, out_vel);
    cKinTree::SetRootAngVel(root_ang_vel, out_vel);
}

// cSceneImitate::CalcRewardImitate (SceneImitate.cpp:7-127); sim_char queries answered kinematically from
// (pose0, vel0), plane ground at height 0, kin_char.GetOriginPos()[1] = ground_h1
double RewardImitate(const Eigen::VectorXd& pose0, const Eigen::VectorXd& vel0, const Eigen::VectorXd& pose1,
                     const Eigen::VectorXd& vel1, double ground_h1, double* out_terms) {
    double pose_w = 0.5, vel_w = 0.05, end_eff_w = 0.15, root_w = 0.2, com_w = 0.1;
    double total_w = pose_w + vel_w + end_eff_w + root_w + com_w;
    pose_w /= total_w; vel_w /= total_w; end_eff_w /= total_w; root_w /= total_w; com_w /= total_w;

    const Eigen::MatrixXd& joint_mat = g.joint_mat;
    const Eigen::MatrixXd& body_defs = g.body_defs;
    int num_joints = g.num_joints;
    const double pose_scale = 2.0 / 15 * num_joints;
    const double vel_scale = 0.1 / 15 * num_joints;
    const double end_eff_scale = 10, root_scale = 5, com_scale = 10, err_scale = 1;

    tMatrix origin_trans = cKinTree::BuildOriginTrans(pose0);
    tMatrix kin_origin_trans = cKinTree::BuildOriginTrans(pose1);

    tVector com0_world, com_vel0_world, com1_world, com_vel1_world;
    cRBDUtil::CalcCoM(joint_mat, body_defs, pose0, vel0, com0_world, com_vel0_world);
    cRBDUtil::CalcCoM(joint_mat, body_defs, pose1, vel1, com1_world, com_vel1_world);

    int root_id = cKinTree::GetRootID();
    tVector root_pos0 = cKinTree::GetRootPos(pose0);
    tVector root_pos1 = cKinTree::GetRootPos(pose1);
    tQuaternion root_rot0 = cKinTree::GetRootRot(pose0);
    tQuaternion root_rot1 = cKinTree::GetRootRot(pose1);
    tVector root_vel0 = cKinTree::GetRootVel(vel0);
    tVector root_vel1 = cKinTree::GetRootVel(vel1);
    tVector root_ang_vel0 = cKinTree::GetRootAngVel(vel0);
    tVector root_ang_vel1 = cKinTree::GetRootAngVel(vel1);

    double pose_err = 0, vel_err = 0, end_eff_err = 0;
    double root_rot_w = g.joint_weights[root_id];
    pose_err += root_rot_w * cKinTree::CalcRootRotErr(joint_mat, pose0, pose1);
    vel_err += root_rot_w * cKinTree::CalcRootAngVelErr(joint_mat, vel0, vel1);

    for (int j = root_id + 1; j < num_joints; ++j) {
        double w = g.joint_weights[j];
        pose_err += w * cKinTree::CalcPoseErr(joint_mat, j, pose0, pose1);
        vel_err += w * cKinTree::CalcVelErr(joint_mat, j, vel0, vel1);
        if (cKinTree::IsEndEffector(joint_mat, j)) {
            tVector pos0 = cKinTree::CalcJointWorldPos(joint_mat, pose0, j);
            tVector pos1 = cKinTree::CalcJointWorldPos(joint_mat, pose1, j);
            double ground_h0 = 0;
            tVector pos_rel0 = pos0 - root_pos0;
            tVector pos_rel1 = pos1 - root_pos1;
            pos_rel0[1] = pos0[1] - ground_h0;
            pos_rel1[1] = pos1[1] - ground_h1;
            pos_rel0 = origin_trans * pos_rel0;
            pos_rel1 = kin_origin_trans * pos_rel1;
            end_eff_err += (pos_rel1 - pos_rel0).squaredNorm();
        }
    }

    root_pos0[1] -= 0;
    root_pos1[1] -= ground_h1;
    double root_pos_err = (root_pos0 - root_pos1).squaredNorm();
    double root_rot_err = cMathUtil::QuatDiffTheta(root_rot0, root_rot1);
    root_rot_err *= root_rot_err;
    double root_vel_err = (root_vel1 - root_vel0).squaredNorm();
    double root_ang_vel_err = (root_ang_vel1 - root_ang_vel0).squaredNorm();
    double root_err = root_pos_err + 0.1 * root_rot_err + 0.01 * root_vel_err + 0.001 * root_ang_vel_err;
    double com_err = 0.1 * (com_vel1_world - com_vel0_world).squaredNorm();

    double terms[5] = {exp(-err_scale * pose_scale * pose_err), exp(-err_scale * vel_scale * vel_err),
                       exp(-err_scale * end_eff_scale * end_eff_err), exp(-err_scale * root_scale * root_err),
                       exp(-err_scale * com_scale * com_err)};
    if (out_terms) std::memcpy(out_terms, terms, sizeof(terms));
    return pose_w * terms[0] + vel_w * terms[1] + end_eff_w * terms[2] + root_w * terms[3] + com_w * terms[4];
}

}  // namespace

extern "C" {

// Loads the character (skeleton + body defs, Character.cpp:35-60 / KinTree.cpp:125-170) and the motion clip
// (KinController.cpp:101-160).  Returns 0 on success.
int dmref_init(const char* char_file, const char* motion_file) {
    g = tState();
    std::ifstream f_stream(char_file);
    Json::Reader reader;
    Json::Value root;
    bool succ = reader.parse(f_stream, root);
    f_stream.close();
    if (!succ || root["Skeleton"].isNull()) return 1;
    std::vector<std::string> names;
    if (!cKinTree::Load(root["Skeleton"], g.joint_mat, names)) return 2;
    if (!cKinTree::LoadBodyDefs(char_file, g.body_defs)) return 3;
    g.num_joints = cKinTree::GetNumJoints(g.joint_mat);
    g.dof = cKinTree::GetNumDof(g.joint_mat);

    // SceneImitate.cpp:300-312
    g.joint_weights = Eigen::VectorXd::Ones(g.num_joints);
    double sum = 0;
    for (int j = 0; j < g.num_joints; ++j) {
        g.joint_weights[j] = cKinTree::GetJointDiffWeight(g.joint_mat, j);
        sum += std::abs(g.joint_weights[j]);
    }
    g.joint_weights /= sum;

    // KinController.cpp:101-117 (the three callbacks bind cKinCharacter members that forward to cKinTree,
    // KinCharacter.cpp:684-704)
    cMotion::tParams params;
    params.mMotionFile = motion_file;
    params.mBlendFunc = [](const cMotion::tFrame* a, const cMotion::tFrame* b, double lerp, cMotion::tFrame* out) {
        cKinTree::LerpPoses(g.joint_mat, *a, *b, lerp, *out);
    };
    params.mVelFunc = [](const cMotion::tFrame* a, const cMotion::tFrame* b, double dt, cMotion::tFrame* out) {
        cKinTree::CalcVel(g.joint_mat, *a, *b, dt, *out);
    };
    params.mPostProcessFunc = [](cMotion::tFrame* out) { cKinTree::PostProcessPose(g.joint_mat, *out); };
    if (!g.motion.Load(params)) return 4;
    if (g.motion.GetNumDof() != g.dof) return 5;

    // KinController.cpp:144-160 (PostProcessMotion: first frame's xz to the origin)
    Eigen::VectorXd frame_beg = g.motion.GetFrame(0);
    tVector root_pos_beg = cKinTree::GetRootPos(frame_beg);
    int num_frames = g.motion.GetNumFrames();
    for (int f = 0; f < num_frames; ++f) {
        Eigen::VectorXd frame = g.motion.GetFrame(f);
        tVector root_pos = cKinTree::GetRootPos(frame);
        root_pos[0] -= root_pos_beg[0];
        root_pos[2] -= root_pos_beg[2];
        cKinTree::SetRootPos(root_pos, frame);
        g.motion.SetFrame(f, frame);
    }
    // KinController.cpp:162-175
    Eigen::VectorXd fb = g.motion.GetFrame(0);
    Eigen::VectorXd fe = g.motion.GetFrame(num_frames - 1);
    g.cycle_root_delta = cKinTree::GetRootPos(fe) - cKinTree::GetRootPos(fb);
    g.cycle_root_delta[1] = 0;
    g.ok = true;
    return 0;
}

int dmref_num_dof() { return g.dof; }
int dmref_num_joints() { return g.num_joints; }
int dmref_num_frames() { return g.motion.GetNumFrames(); }
double dmref_duration() { return g.motion.GetDuration(); }
int dmref_loop() { return g.motion.EnableLoop() ? 1 : 0; }
void dmref_joint_weights(double* out) { FromVec(g.joint_weights, out); }
int dmref_param_offset(int j) { return cKinTree::GetParamOffset(g.joint_mat, j); }
int dmref_param_size(int j) { return cKinTree::GetParamSize(g.joint_mat, j); }

// the clip table after loading: frames [n][dof], frame velocities [n][dof], frame start times [n]
void dmref_clip_table(double* frames, double* vels, double* times) {
    int n = g.motion.GetNumFrames();
    for (int f = 0; f < n; ++f) {
        FromVec(g.motion.GetFrame(f), frames + f * g.dof);
        FromVec(g.motion.GetFrameVel(f), vels + f * g.dof);
        times[f] = g.motion.GetFrameTime(f);
    }
}

void dmref_kin_pose_vel(double time, const double* origin3, double* out_pose, double* out_vel) {
    Eigen::VectorXd pose, vel;
    tVector origin(origin3 ? origin3[0] : 0, origin3 ? origin3[1] : 0, origin3 ? origin3[2] : 0, 0);
    KinPose(time, origin, pose);
    KinVel(time, vel);
    FromVec(pose, out_pose);
    FromVec(vel, out_vel);
}

double dmref_pose_err(int j, const double* pose0, const double* pose1) {
    Eigen::VectorXd p0 = ToVec(pose0, g.dof), p1 = ToVec(pose1, g.dof);
    return j == 0 ? cKinTree::CalcRootRotErr(g.joint_mat, p0, p1) : cKinTree::CalcPoseErr(g.joint_mat, j, p0, p1);
}
double dmref_vel_err(int j, const double* vel0, const double* vel1) {
    Eigen::VectorXd v0 = ToVec(vel0, g.dof), v1 = ToVec(vel1, g.dof);
    return j == 0 ? cKinTree::CalcRootAngVelErr(g.joint_mat, v0, v1) : cKinTree::CalcVelErr(g.joint_mat, j, v0, v1);
}
void dmref_joint_world_pos(const double* pose, int j, double* out3) {
    tVector p = cKinTree::CalcJointWorldPos(g.joint_mat, ToVec(pose, g.dof), j);
    out3[0] = p[0]; out3[1] = p[1]; out3[2] = p[2];
}
void dmref_joint_world_trans(const double* pose, int j, double* out16) {
    FromMat4(cKinTree::JointWorldTrans(g.joint_mat, ToVec(pose, g.dof), j), out16);
}
double dmref_heading(const double* pose) { return cKinTree::CalcHeading(ToVec(pose, g.dof)); }
void dmref_origin_trans(const double* pose, double* out16) { FromMat4(cKinTree::BuildOriginTrans(ToVec(pose, g.dof)), out16); }
void dmref_com(const double* pose, const double* vel, double* com3, double* com_vel3) {
    tVector com, com_vel;
    cRBDUtil::CalcCoM(g.joint_mat, g.body_defs, ToVec(pose, g.dof), ToVec(vel, g.dof), com, com_vel);
    for (int i = 0; i < 3; ++i) { com3[i] = com[i]; com_vel3[i] = com_vel[i]; }
}
void dmref_lerp_poses(const double* pose0, const double* pose1, double lerp, double* out) {
    Eigen::VectorXd r;
    cKinTree::LerpPoses(g.joint_mat, ToVec(pose0, g.dof), ToVec(pose1, g.dof), lerp, r);
    FromVec(r, out);
}
void dmref_calc_vel(const double* pose0, const double* pose1, double dt, double* out) {
    Eigen::VectorXd r;
    cKinTree::CalcVel(g.joint_mat, ToVec(pose0, g.dof), ToVec(pose1, g.dof), dt, r);
    FromVec(r, out);
}
double dmref_quat_theta(const double* q_wxyz) {
    return cMathUtil::QuatTheta(tQuaternion(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]));
}
void dmref_quat_rot_vec(const double* q_wxyz, const double* v3, double* out3) {
    tVector r = cMathUtil::QuatRotVec(tQuaternion(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]), tVector(v3[0], v3[1], v3[2], 0));
    out3[0] = r[0]; out3[1] = r[1]; out3[2] = r[2];
}
void dmref_normal_tangent(const double* q_wxyz, double* norm3, double* tan3) {
    tVector n, t;
    cMathUtil::CalcNormalTangent(tQuaternion(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]), n, t);
    for (int i = 0; i < 3; ++i) { norm3[i] = n[i]; tan3[i] = t[i]; }
}

double dmref_reward(const double* pose0, const double* vel0, const double* pose1, const double* vel1,
                    double ground_h1, double* out_terms5) {
    return RewardImitate(ToVec(pose0, g.dof), ToVec(vel0, g.dof), ToVec(pose1, g.dof), ToVec(vel1, g.dof),
                         ground_h1, out_terms5);
}

// n poses against the clip at their own times; origin may be NULL (kinematic character at the world origin)
void dmref_reward_batch(int n, const double* pose, const double* vel, const double* time, const double* origin,
                        double* out_reward, double* out_terms) {
    for (int e = 0; e < n; ++e) {
        Eigen::VectorXd p1, v1;
        tVector org(origin ? origin[3 * e] : 0, origin ? origin[3 * e + 1] : 0, origin ? origin[3 * e + 2] : 0, 0);
        KinPose(time[e], org, p1);
        KinVel(time[e], v1);
        out_reward[e] = RewardImitate(ToVec(pose + e * g.dof, g.dof), ToVec(vel + e * g.dof, g.dof), p1, v1, org[1],
                                      out_terms ? out_terms + 5 * e : nullptr);
    }
}

// cCtController::BuildStatePose + BuildStateVel (CtController.cpp:378-495), plane ground at height 0, bodies
// placed kinematically.  out = [root height, 15 x (pos3, normal3, tangent3), 15 x (lin vel3, ang vel3)]
void dmref_record_state(const double* pose_in, const double* vel_in, int record_all_world, int record_world_root_pos,
                        int record_world_root_rot, double vel_scale, double* out) {
    Eigen::VectorXd pose = ToVec(pose_in, g.dof), vel = ToVec(vel_in, g.dof);
    tMatrix origin_trans = cKinTree::BuildOriginTrans(pose);
    tQuaternion origin_quat = cMathUtil::RotMatToQuaternion(origin_trans);
    tVector root_pos = cKinTree::GetRootPos(pose);
    double ground_h = 0;
    tVector root_pos_rel = root_pos;
    root_pos_rel[1] -= ground_h;
    root_pos_rel[3] = 1;
    root_pos_rel = origin_trans * root_pos_rel;
    root_pos_rel[3] = 0;

    const int pos_dim = 3, rot_dim = 6, vel_dim = 3, ang_vel_dim = 3;
    int num_parts = g.num_joints;
    int root_id = cKinTree::GetRootID();
    Eigen::VectorXd out_pose = Eigen::VectorXd::Zero(1 + num_parts * (pos_dim + rot_dim));
    Eigen::VectorXd out_vel = Eigen::VectorXd::Zero(num_parts * (vel_dim + ang_vel_dim));
    out_pose[0] = root_pos_rel[1];
    for (int i = 0; i < num_parts; ++i) {
        if (!cKinTree::IsValidBody(g.body_defs, i)) continue;
        tMatrix body_world = cKinTree::BodyWorldTrans(g.joint_mat, g.body_defs, pose, i);
        tVector curr_pos = cKinTree::CalcBodyPartPos(g.joint_mat, g.body_defs, pose, i);
        curr_pos[1] -= ground_h;
        if (!record_all_world) {
            if (!record_world_root_pos || i != root_id) {
                curr_pos[3] = 1;
                curr_pos = origin_trans * curr_pos;
                curr_pos -= root_pos_rel;
                curr_pos[3] = 0;
            }
        }
        int pos_idx = (pos_dim + rot_dim) * i + 1;
        out_pose.segment(pos_idx, pos_dim) = curr_pos.segment(0, pos_dim);

        tQuaternion curr_quat = cMathUtil::RotMatToQuaternion(body_world);
        if (!record_all_world) {
            if (!record_world_root_rot || i != root_id) curr_quat = origin_quat * curr_quat;
        }
        tVector curr_rot_norm, curr_rot_tan;
        cMathUtil::CalcNormalTangent(curr_quat, curr_rot_norm, curr_rot_tan);
        int rot_idx = (pos_dim + rot_dim) * i + 1 + pos_dim;
        out_pose.segment(rot_idx, rot_dim / 2) = curr_rot_norm.segment(0, rot_dim / 2);
        out_pose.segment(rot_idx + rot_dim / 2, rot_dim / 2) = curr_rot_tan.segment(0, rot_dim / 2);

        tVector curr_vel = cKinTree::CalcBodyPartVel(g.joint_mat, g.body_defs, pose, vel, i);
        tVector curr_ang_vel = cKinTree::CalcJointWorldAngularVel(g.joint_mat, pose, vel, i);
        if (!record_all_world) {
            if (!record_world_root_rot || i != root_id) {
                curr_vel = origin_trans * curr_vel;
                curr_ang_vel = origin_trans * curr_ang_vel;
            }
        }
        int vel_idx = (vel_dim + ang_vel_dim) * i;
        out_vel.segment(vel_idx, vel_dim) = curr_vel.segment(0, vel_dim) * vel_scale;
        out_vel.segment(vel_idx + vel_dim, ang_vel_dim) = curr_ang_vel.segment(0, ang_vel_dim) * vel_scale;
    }
    FromVec(out_pose, out);
    FromVec(out_vel, out + out_pose.size());
}

// cKinCharacter::AddNoise over AddNoisePoseVel and RandomRotatePoseVel (anim/KinCharacter.cpp:340-532) with
// cCharacter::RotateRoot (anim/Character.cpp:210-216): KinCharacter.cpp itself needs the OpenGL headers (Character.h
// includes render/DrawMesh.h) and cannot be compiled here, so the loop is restated around the reference's own
// cMathUtil::{AxisAngleToQuaternion, EulerToQuaternion, VecToQuat, QuatToVec} and cKinTree::{Get/SetRootRot,
// Get/SetRootVel, Get/SetRootAngVel, GetParamOffset/Size, GetJointType, PostProcessPose}.  The reference draws from its
// global generator (cMathUtil::RandDouble); here every draw is an INPUT, consumed in the reference's order:
//   u_pose[dof], u_vel[dof] in [0, 1)  - AddNoisePoseVel, scaled to [noise_min, noise_max)
//   r[...] in [-1, 1)                  - RandomRotatePoseVel, scaled by `radian`: root yaw, then per joint 1 (revolute) or
//                                        3 (spherical) values for the pose, then (vel_noise) 3 for the root's angular
//                                        velocity and again 1 / 3 per joint; joints that get no noise consume none.
// Returns the number of r values consumed.
int dmref_reset_noise(const double* pose_in, const double* vel_in, int noise_bef_rot, double noise_min, double noise_max,
                      double radian, int rot_vel_w_pose, int vel_noise, double interp, int knee_rot,
                      const double* u_pose, const double* u_vel, const double* r, double* pose_out, double* vel_out) {
    Eigen::VectorXd mPose = ToVec(pose_in, g.dof), mVel = ToVec(vel_in, g.dof);
    auto add_noise_pose_vel = [&]() {
        if (noise_min == 0 && noise_max == 0) return;
        for (int i = 0; i < g.dof; ++i) mPose[i] += noise_min + (noise_max - noise_min) * u_pose[i];
        for (int i = 0; i < g.dof; ++i) mVel[i] += noise_min + (noise_max - noise_min) * u_vel[i];
    };
    int used = 0;
    auto random_rotate = [&]() {
        if (radian == 0) return;
        double range = radian;
        auto draw = [&]() { return range * r[used++]; };
        tQuaternion root_rotate = cMathUtil::AxisAngleToQuaternion(tVector(0, 1, 0, 0), draw());
        tQuaternion root_rot = cKinTree::GetRootRot(mPose);
        root_rot = root_rotate * root_rot;
        root_rot.normalize();
        cKinTree::SetRootRot(root_rot, mPose);

        tVector vel = cKinTree::GetRootVel(mVel);
        vel = interp * vel;
        cKinTree::SetRootVel(vel, mVel);
        tVector ang_vel = cKinTree::GetRootAngVel(mVel);
        ang_vel = interp * ang_vel;
        cKinTree::SetRootAngVel(ang_vel, mVel);
        for (int j = 1; j < g.num_joints; ++j) {
            int param_offset = cKinTree::GetParamOffset(g.joint_mat, j);
            int param_size = cKinTree::GetParamSize(g.joint_mat, j);
            tVector v = tVector::Zero();
            v.segment(0, param_size) = mVel.segment(param_offset, param_size);
            v = interp * v;
            mVel.segment(param_offset, param_size) = v.segment(0, param_size);
        }
        for (int j = 1; j < g.num_joints; ++j) {
            int param_offset = cKinTree::GetParamOffset(g.joint_mat, j);
            int param_size = cKinTree::GetParamSize(g.joint_mat, j);
            cKinTree::eJointType joint_type = cKinTree::GetJointType(g.joint_mat, j);
            if (joint_type == cKinTree::eJointTypeRevolute) {
                if (!(j == 4 || j == 10) || knee_rot) mPose(param_offset) = mPose(param_offset) + draw();
            } else if (joint_type == cKinTree::eJointTypeSpherical) {
                if (j != 3 && j != 5 && j != 9 && j != 11) {
                    double rand_psi = draw();
                    double rand_theta = draw();
                    double rand_phi = draw();
                    tQuaternion rand_rotate = cMathUtil::EulerToQuaternion(tVector(rand_psi, rand_theta, rand_phi, 0));
                    tVector seg = tVector::Zero();
                    seg.segment(0, param_size) = mPose.segment(param_offset, param_size);
                    tQuaternion rot = cMathUtil::VecToQuat(seg);
                    rot = rand_rotate * rot;
                    mPose.segment(param_offset, param_size) = cMathUtil::QuatToVec(rot).segment(0, param_size);
                    if (rot_vel_w_pose) {
                        tVector vseg = tVector::Zero();
                        vseg.segment(0, param_size) = mVel.segment(param_offset, param_size);
                        tQuaternion vq = cMathUtil::VecToQuat(vseg);
                        vq = rand_rotate * vq;
                        mVel.segment(param_offset, param_size) = cMathUtil::QuatToVec(vq).segment(0, param_size);
                    }
                }
            }
        }
        if (vel_noise) {
            tQuaternion av = cMathUtil::VecToQuat(cKinTree::GetRootAngVel(mVel));
            double rand_psi = draw();
            double rand_theta = draw();
            double rand_phi = draw();
            tQuaternion rand_rotate = cMathUtil::EulerToQuaternion(tVector(rand_psi, rand_theta, rand_phi, 0));
            av = rand_rotate * av;
            cKinTree::SetRootAngVel(cMathUtil::QuatToVec(av), mVel);
            for (int j = 1; j < g.num_joints; ++j) {
                int param_offset = cKinTree::GetParamOffset(g.joint_mat, j);
                int param_size = cKinTree::GetParamSize(g.joint_mat, j);
                cKinTree::eJointType joint_type = cKinTree::GetJointType(g.joint_mat, j);
                if (joint_type == cKinTree::eJointTypeRevolute) {
                    if (!(j == 4 || j != 10) || knee_rot) mVel(param_offset) = mVel(param_offset) + draw();
                } else if (joint_type == cKinTree::eJointTypeSpherical) {
                    if (j != 3 && j != 5 && j != 9 && j != 11) {
                        double p0 = draw();
                        double p1 = draw();
                        double p2 = draw();
                        tQuaternion rr = cMathUtil::EulerToQuaternion(tVector(p0, p1, p2, 0));
                        tVector vseg = tVector::Zero();
                        vseg.segment(0, param_size) = mVel.segment(param_offset, param_size);
                        tQuaternion vq = cMathUtil::VecToQuat(vseg);
                        vq = rr * vq;
                        mVel.segment(param_offset, param_size) = cMathUtil::QuatToVec(vq).segment(0, param_size);
                    }
                }
            }
        }
        cKinTree::PostProcessPose(g.joint_mat, mPose);
    };
    if (noise_bef_rot) {
        add_noise_pose_vel();
        random_rotate();
    } else {
        random_rotate();
        add_noise_pose_vel();
    }
    FromVec(mPose, pose_out);
    FromVec(mVel, vel_out);
    return used;
}

}  // extern "C"
