"""Packs the reference's humanoid3d_spinkick motion clip (mocap data, not code) into
amp_extensions_b200/data/humanoid3d_spinkick.npz.  Run in the build container only."""
import json
import os

import numpy as np

REF = os.environ.get("SIMSTEP_REFERENCE", "/root/reference")
src = os.path.join(REF, "deepmimic/deepmimic/data/motions/humanoid3d_spinkick.txt")
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "amp_extensions_b200", "data",
                   "humanoid3d_spinkick.npz")
with open(src) as f:
    d = json.load(f)
raw = np.array(d["Frames"], dtype=np.float64)
np.savez_compressed(dst, frames_raw=raw, loop=np.array(d["Loop"]))
print("wrote", os.path.normpath(dst), raw.shape, d["Loop"])
