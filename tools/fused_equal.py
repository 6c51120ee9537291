"""Fused forward (SIMSTEP_FUSED_LAYERS=1 in this process) against the per-layer path in a child process: the two
paths run the same tiles with the same arithmetic, so forward outputs must be bit-identical."""
import os, subprocess, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import helpers as H
from tests.test_parity_gpu import make_engine

def run(path):
    c = H.ns_case()
    g = torch.Generator().manual_seed(3)
    out = {}
    for E in (1, 300, 5000, 40000):
        s, a = torch.randn(E, 226, generator=g).cuda(), torch.randn(E, 28, generator=g).cuda()
        eng = make_engine(c, "fp16")
        f = eng.forward(s, a)
        d = eng.discrepancy(s, a)
        out[E] = (f.cpu(), d.cpu())
    torch.save(out, path)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        run("/tmp/fused.pt")
        env = dict(os.environ)
        env.pop("SIMSTEP_FUSED_LAYERS", None)
        subprocess.run([sys.executable, __file__, "/tmp/base.pt"], check=True, env=env)
        a, b = torch.load("/tmp/fused.pt"), torch.load("/tmp/base.pt")
        for E in a:
            print(E, "forward equal", torch.equal(a[E][0], b[E][0]), "disc equal", torch.equal(a[E][1], b[E][1]),
                  "max abs diff", float((a[E][0] - b[E][0]).abs().max()))
        assert all(torch.equal(a[E][0], b[E][0]) for E in a)
