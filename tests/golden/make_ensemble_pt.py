"""Writes tests/golden/ref_ensemble_tiny_dense.pt with the REFERENCE's own DynamicsEnsemble.save_ensemble
(milo/milo/dynamics.py:110-116) for the "tiny_dense" case of make_golden.py — the on-disk format the drop-in
must read and write.  Run in the build container only (needs /root/reference):

    python tests/golden/make_ensemble_pt.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_ensemble_tiny_dense.pt")


def main():
    ref = mg.load_reference()
    S, A, N, hidden, dense, act = 20, 6, 3, [32, 24], True, "relu"
    s, a, s2 = mg.synth_dataset(512, S, A, seed=0)
    ds = ref["datasets"].AmpDataset(s, a, s2)
    ens = mg.build_ensemble(ref, S, A, N, hidden, dense, act, ds)
    ens.save_ensemble(OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
