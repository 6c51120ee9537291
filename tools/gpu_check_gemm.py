"""GPU bring-up check of the tcgen05 GEMM and the raw ensemble forward (run under gpurun).

Prints one line per case; exits non-zero on the first numerical failure.  Independent of the Python
shims: talks to libsimstep.so through ctypes only, and uses torch on the GPU as the fp64 checker.
"""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from amp_extensions_b200 import _lib  # noqa: E402


def round_operand(x, prec):
    if prec == "tf32":
        i = x.contiguous().view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32)
    if prec == "fp16":
        return x.clamp(-65504, 65504).half().float()
    return x.bfloat16().float()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def check_gemm(lib, prec, groups, m, n, k, seed=0, with_bias=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(groups, m, k, device="cuda", generator=g)
    b = torch.randn(groups, n, k, device="cuda", generator=g) / (k ** 0.5)
    bias = torch.randn(groups, n, device="cuda", generator=g) if with_bias else None
    d = torch.full((groups, m, n), float("nan"), device="cuda")
    rc = lib.simstep_debug_gemm(_lib.PREC[prec], groups, m, n, k, ptr(a), ptr(b), ptr(bias), ptr(d), None)
    if rc != 0:
        print("FAIL rc", rc, lib.simstep_last_error(None))
        return False
    torch.cuda.synchronize()
    ar, br = round_operand(a, prec).double(), round_operand(b, prec).double()
    ref = torch.einsum("gmk,gnk->gmn", ar, br)
    if bias is not None:
        ref = ref + bias.double()[:, None, :]
    err = (d.double() - ref).abs()
    scale = ref.abs().max().item()
    mx = err.max().item()
    ok = bool(torch.isfinite(d).all().item()) and mx <= 2e-5 * max(scale, 1.0) * max(1.0, (k / 256) ** 0.5)
    print(f"gemm {prec} g={groups} m={m} n={n} k={k}: max_abs_err={mx:.3e} ref_max={scale:.3e} {'ok' if ok else 'FAIL'}")
    if not ok:
        e = err[0]
        print("  nan count", int(torch.isnan(d).sum().item()), "of", d.numel())
        # error map by 32-row x 32-col blocks of the first tile
        mm, nn = min(m, 128), min(n, 256)
        blk = e[:mm, :nn]
        for r0 in range(0, mm, 32):
            print("  rows %3d.. " % r0 + " ".join(f"{blk[r0:r0+32, c0:c0+32].max().item():8.2e}" for c0 in range(0, nn, 32)))
        print("  first row got ", d[0, 0, :8].tolist())
        print("  first row want", ref[0, 0, :8].tolist())
    return ok


def make_mlp(N, S, A, hidden, seed):
    g = torch.Generator().manual_seed(seed)
    sizes = [S + A] + hidden + [S]
    Ws, bs = [], []
    for m in range(N):
        w_m, b_m = [], []
        for i in range(len(sizes) - 1):
            fan_in = sum(sizes[: i + 1])
            bound = 1.0 / fan_in ** 0.5
            w_m.append(((torch.rand(sizes[i + 1], fan_in, generator=g) * 2 - 1) * bound).contiguous())
            b_m.append(((torch.rand(sizes[i + 1], generator=g) * 2 - 1) * bound).contiguous())
        Ws.append(w_m)
        bs.append(b_m)
    return Ws, bs


def torch_forward(Ws, bs, tf, s, a):
    x = torch.cat([(s - tf[0]) / tf[1], (a - tf[2]) / tf[3]], 1)
    outs = []
    for w_m, b_m in zip(Ws, bs):
        inp = x
        for i in range(len(w_m) - 1):
            out = torch.relu(inp @ w_m[i].T + b_m[i])
            inp = torch.cat([inp, out], 1)
        y = inp @ w_m[-1].T + b_m[-1]
        outs.append(y * tf[5] + tf[4])
    return torch.stack(outs, 0)


def check_forward(lib, prec, N, hidden, E, bench_iters=0):
    S, A = 226, 28
    Ws, bs = make_mlp(N, S, A, hidden, seed=100)
    g = torch.Generator().manual_seed(1)
    tf = [torch.randn(S, generator=g) * 0.1, torch.rand(S, generator=g) + 0.5, torch.randn(A, generator=g) * 0.1,
          torch.rand(A, generator=g) + 0.5, torch.randn(S, generator=g) * 0.01, torch.rand(S, generator=g) * 0.1 + 0.01]
    cfg = _lib.SimstepConfig()
    cfg.abi_version = _lib.ABI_VERSION
    cfg.state_dim, cfg.action_dim, cfg.n_models, cfg.n_hidden = S, A, N, len(hidden)
    for i, hsz in enumerate(hidden):
        cfg.hidden[i] = hsz
    cfg.dense_connect, cfg.activation, cfg.transform, cfg.precision = 1, 0, 1, _lib.PREC[prec]
    h = C.c_void_p()
    _lib.check(lib.simstep_create(C.byref(cfg), C.byref(h)))
    nl = len(hidden) + 1
    wp = (C.c_void_p * (N * nl))(*[Ws[m][l].data_ptr() for m in range(N) for l in range(nl)])
    bp = (C.c_void_p * (N * nl))(*[bs[m][l].data_ptr() for m in range(N) for l in range(nl)])
    tp = (C.c_void_p * 6)(*[t.data_ptr() for t in tf])
    _lib.check(lib.simstep_load_ensemble(h, wp, bp, tp), h)
    s = torch.randn(E, S, generator=g).cuda()
    a = torch.randn(E, A, generator=g).cuda()
    delta = torch.full((N, E, S), float("nan"), device="cuda")
    _lib.check(lib.simstep_forward(h, ptr(s), ptr(a), E, ptr(delta), None), h)
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    Wd = [[w.cuda().double() for w in wm] for wm in Ws]
    bd = [[b.cuda().double() for b in bm] for bm in bs]
    ref = torch_forward(Wd, bd, [t.cuda().double() for t in tf], s.double(), a.double())
    err = (delta.double() - ref)
    rel_l2 = (err.norm() / ref.norm()).item()
    mx = err.abs().max().item()
    ok = bool(torch.isfinite(delta).all().item()) and rel_l2 < (5e-3 if prec == "bf16" else 1e-3)
    print(f"forward {prec} N={N} hidden={hidden} E={E}: rel_l2={rel_l2:.3e} max_abs={mx:.3e} ref_max={ref.abs().max().item():.3e} {'ok' if ok else 'FAIL'}")
    if bench_iters:
        disc = torch.empty(E, device="cuda")
        for _ in range(3):
            _lib.check(lib.simstep_discrepancy(h, ptr(s), ptr(a), E, ptr(disc), None), h)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(bench_iters):
            _lib.check(lib.simstep_discrepancy(h, ptr(s), ptr(a), E, ptr(disc), None), h)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / bench_iters
        sizes = [S + A] + hidden + [S]
        flop = 0
        for i in range(len(sizes) - 1):
            flop += 2 * sum(sizes[: i + 1]) * sizes[i + 1]
        tfl = flop * N * E / (ms * 1e-3) / 1e12
        print(f"  bench {prec}: {ms:.3f} ms per pass, {E / (ms * 1e-3):.3e} env-steps/s, {tfl:.1f} TFLOP/s algorithmic")
    _lib.check(lib.simstep_destroy(h))
    return ok


def main():
    lib = _lib.load()
    print("device", torch.cuda.get_device_name(0))
    ok = True
    precs = ["tf32", "fp16", "bf16"]
    for prec in precs:
        ok &= check_gemm(lib, prec, 1, 128, 256, 64)
        ok &= check_gemm(lib, prec, 1, 128, 256, 256)
        if not ok:
            break
        ok &= check_gemm(lib, prec, 1, 256, 512, 512)
        ok &= check_gemm(lib, prec, 2, 300, 500, 700, with_bias=True)
        ok &= check_gemm(lib, prec, 4, 1000, 226, 2302)
        ok &= check_gemm(lib, prec, 3, 40000, 512, 1278)
    if ok:
        for prec in precs:
            ok &= check_forward(lib, prec, 4, [512] * 4, 1000)
            ok &= check_forward(lib, prec, 4, [512] * 4, 40000, bench_iters=10)
        ok &= check_forward(lib, "tf32", 8, [1024] * 4, 20000, bench_iters=5)
        ok &= check_forward(lib, "tf32", 2, [64, 96], 300)
    print("launches", _lib.launch_count())
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    t0 = time.time()
    main()
