"""BASELINE.json configs[2]: DeepMimic/AMP imitation reward only, 1M synthetic humanoid3d poses vs the spinkick
clip.  Prints one JSON line (poses/s and achieved algorithmic HBM GB/s = 352 B/pose).

    python tools/bench_imitation.py [--poses 1048576] [--iters 50] [--origin] [--terms]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_poses(imit, E, device, seed=3):
    """pose = clip(t) perturbed, vel = clipvel(t) + N(0, .5), t ~ U(0, duration)."""
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.rand(E, device=device, generator=g) * float(imit.clip.duration)
    pose, vel = imit.sample(t)
    pose = pose + 0.05 * torch.randn(pose.shape, device=device, generator=g)
    vel = vel + 0.5 * torch.randn(vel.shape, device=device, generator=g)
    ch = imit.character
    for j, jt in enumerate(ch.joint_type):
        o = ch.param_offset[j] + (3 if j == 0 else 0)
        if j == 0 or jt == 1:  # quaternion segments stay unit length
            q = pose[:, o:o + 4]
            pose[:, o:o + 4] = q / q.norm(dim=1, keepdim=True)
    return pose.contiguous(), vel.contiguous(), t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=1 << 20)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--origin", action="store_true")
    ap.add_argument("--terms", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    from amp_extensions_b200 import ImitationReward
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    imit = ImitationReward(device=dev)
    E = args.poses
    ring = 4  # 4 x 360 MB of inputs > 126 MB L2
    sets = [synth_poses(imit, E, dev, seed=3 + i) for i in range(ring)]
    origin = torch.zeros(E, 3, device=dev) if args.origin else None
    for i in range(args.warmup):
        p, v, t = sets[i % ring]
        imit.reward(p, v, t, origin, want_terms=args.terms)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.iters):
        p, v, t = sets[i % ring]
        r = imit.reward(p, v, t, origin, want_terms=args.terms)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    bytes_per_pose = 352 + (12 if args.origin else 0) + (20 if args.terms else 0)
    peak = 6543.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    gbs = bytes_per_pose * E / (ms * 1e-3) / 1e9
    rr = r[0] if isinstance(r, tuple) else r
    cpu = None
    if not args.skip_cpu and args.iters >= 10:
        # CPU baseline on a bounded sample: the reference's own compiled kinematics code (oracle/_ref/libdmref.so, built
        # by oracle/ref_build.py from the reference sources; "reference") when it was shipped, else the float64 numpy
        # restatement of CalcRewardImitate ("port"); one core either way (the reference's reward is single-threaded)
        import ctypes
        import time
        import numpy as np
        from oracle import imitation_oracle as io
        from oracle import ref_build
        p, v, t = sets[0]
        lib = ref_build.load()
        if lib is not None:
            n = 4096
            pp, vv, tt = (np.ascontiguousarray(x[:n].double().cpu().numpy()) for x in (p, v, t))
            out = np.zeros(n)
            PD = ctypes.POINTER(ctypes.c_double)
            call = lambda: lib.dmref_reward_batch(n, pp.ctypes.data_as(PD), vv.ctypes.data_as(PD), tt.ctypes.data_as(PD),
                                                  None, out.ctypes.data_as(PD), None)
            kind, what = "reference", "cSceneImitate::CalcRewardImitate over the reference's compiled cKinTree / cRBDUtil / cMotion (oracle/_ref/libdmref.so)"
        else:
            n = 64
            clip = io.Clip(np.load(os.path.join(ROOT, "amp_extensions_b200", "data", "humanoid3d_spinkick.npz"))["frames_raw"],
                           io.HUMANOID3D, "wrap")
            pp, vv, tt = p[:n].double().cpu().numpy(), v[:n].double().cpu().numpy(), t[:n].double().cpu().numpy()
            call = lambda: io.imitation_reward_batch(io.HUMANOID3D, clip, pp, vv, tt)
            kind, what = "port", "numpy float64 restatement of CalcRewardImitate"
        call()
        t0, reps = time.perf_counter(), 0
        while time.perf_counter() - t0 < 8.0:
            call()
            reps += 1
        cpu = {"value": n * reps / (time.perf_counter() - t0), "unit": "poses/s", "cores": 1, "kind": kind,
               "sample": f"{reps} x {n} poses, {what}"}
    print(json.dumps({"workload": f"imitation reward, {E} poses vs spinkick clip", "ms_per_launch": ms,
                      "poses_per_s": E / (ms * 1e-3), "bytes_per_pose": bytes_per_pose, "achieved_gbs": gbs,
                      "hbm_peak_gbs": peak, "frac": gbs / peak, "reward_mean": float(rr.mean()),
                      "fp32_roofline": {"flop_per_pose": 4104, "peak_tflops": 148 * 128 * 2 * 1.87e9 / 1e12,
                                        "achieved_tflops": 4104 * E / (ms * 1e-3) / 1e12,
                                        "note": "FLOP per pose counted from the ncu source page (DESIGN.md section 5)"},
                      "cpu_baseline": cpu}))


if __name__ == "__main__":
    main()
