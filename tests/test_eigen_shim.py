"""CPU: the Eigen stand-in (oracle/eigen_shim) checked on its own against numpy / scipy.

oracle/ref_build.py compiles the reference's DeepMimicCore kinematics sources against these headers because Eigen
3.3.7 is not in the image; the float64 restatement agreeing with that build to 1e-14 is one check of the stand-in,
this file is an independent one: dense products, block / row / column / segment views (including the
column-into-row assignment Eigen transposes implicitly), the comma initialiser's row-major order, column-major
data(), and the quaternion formulas (Hamilton product, q*v, slerp with its near-parallel linear branch and its
negative-dot flip, FromTwoVectors) against scipy.spatial.transform.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial.transform import Rotation, Slerp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "eigen_shim")
OUT = os.path.join(ROOT, "oracle", "_ref")
PD = ctypes.POINTER(ctypes.c_double)


def P(a):
    return a.ctypes.data_as(PD)


@pytest.fixture(scope="module")
def lib():
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libshimtest.so")
    src = os.path.join(SHIM, "selftest.cpp")
    deps = [src, os.path.join(SHIM, "Eigen", "Core")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++14", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-I", SHIM, src, "-o", so],
                       check=True)
    handle = ctypes.CDLL(so)
    handle.shim_quat_slerp.argtypes = [PD, PD, ctypes.c_double, PD]
    handle.shim_matmul.argtypes = [PD, PD, ctypes.c_int, ctypes.c_int, ctypes.c_int, PD]
    handle.shim_solve.argtypes = [PD, PD, ctypes.c_int, PD]
    return handle


def wxyz(r):
    q = r.as_quat()  # scipy: (x, y, z, w)
    return np.array([q[3], q[0], q[1], q[2]])


def test_dense_product(lib):
    rng = np.random.default_rng(0)
    for r, k, c in ((4, 4, 4), (6, 6, 1), (3, 7, 5), (1, 9, 1)):
        a, b = rng.normal(size=(r, k)), rng.normal(size=(k, c))
        out = np.zeros((r, c))
        lib.shim_matmul(P(a), P(b), r, k, c, P(out))
        np.testing.assert_allclose(out, a @ b, atol=1e-14)


def test_views_initialiser_and_storage_order(lib):
    out = np.zeros(24)
    lib.shim_views(P(out))
    m = np.arange(1.0, 17.0).reshape(4, 4)              # row-major fill of the comma initialiser
    m[0:3, 0:3] = 2.0 * np.eye(3) + m[1:4, 1:4].copy()  # overlapping source block: evaluated before the write
    m[3, :] = m[:, 3].copy()                             # column -> row
    v = np.zeros(6)
    v[1:4] = 0.5 * m[0:3, 0]
    v[4:6] = [7, 8]
    t = m.T.copy()
    m[2, 1] += t[0, 3]
    np.testing.assert_allclose(out[:16].reshape(4, 4), m, atol=0)
    np.testing.assert_allclose(out[16:22], v, atol=0)
    a, b = np.array([1, 2, 3, 9.0]), np.array([-2, 0.5, 4, 9.0])
    c = np.append(np.cross(a[:3], b[:3]), 0.0)           # cross3: first three entries, w = 0
    assert out[22] == pytest.approx(c.sum(), abs=1e-14)
    expect = ((np.maximum(a, b) - np.minimum(a, b)) ** 2).sum() + np.linalg.norm(a[:3]) + m[1, 0]   # data()[1] = m(1,0)
    assert out[23] == pytest.approx(expect, abs=1e-13)


def test_quaternion_algebra_against_scipy(lib):
    rng = np.random.default_rng(1)
    for _ in range(100):
        ra, rb = Rotation.random(random_state=rng.integers(1 << 30)), Rotation.random(random_state=rng.integers(1 << 30))
        a, b = wxyz(ra), wxyz(rb)
        out = np.zeros(4)
        lib.shim_quat_mul(P(a), P(b), P(out))
        ref = wxyz(ra * rb)
        assert min(np.abs(out - ref).max(), np.abs(out + ref).max()) < 1e-14      # same rotation, either sign
        v, rv = rng.normal(size=3), np.zeros(3)
        lib.shim_quat_rotate(P(a), P(v), P(rv))
        np.testing.assert_allclose(rv, ra.apply(v), atol=1e-14)
        s = a * rng.uniform(0.5, 2.0)                    # not unit: conjugate / inverse / normalized differ
        misc = np.zeros(12)
        lib.shim_quat_misc(P(s), P(misc))
        np.testing.assert_allclose(misc[0:4], s * [1, -1, -1, -1], atol=0)
        np.testing.assert_allclose(misc[4:8], s * [1, -1, -1, -1] / (s @ s), atol=1e-15)
        np.testing.assert_allclose(misc[8:12], s / np.linalg.norm(s), atol=1e-15)


def test_slerp_against_scipy_and_its_branches(lib):
    rng = np.random.default_rng(2)
    for _ in range(100):
        ra, rb = Rotation.random(random_state=rng.integers(1 << 30)), Rotation.random(random_state=rng.integers(1 << 30))
        t = float(rng.uniform(0, 1))
        a, b = wxyz(ra), wxyz(rb)
        out = np.zeros(4)
        lib.shim_quat_slerp(P(a), P(b), t, P(out))
        ref = Slerp([0, 1], Rotation.concatenate([ra, rb]))([t])[0]
        assert (Rotation.from_quat([out[1], out[2], out[3], out[0]]) * ref.inv()).magnitude() < 1e-12
        assert abs(np.linalg.norm(out) - 1) < 1e-14
        # opposite sign of one end: the same rotation, so the same interpolated rotation (the d < 0 flip)
        out2 = np.zeros(4)
        lib.shim_quat_slerp(P(a), P(-b), t, P(out2))
        assert (Rotation.from_quat([out2[1], out2[2], out2[3], out2[0]]) * ref.inv()).magnitude() < 1e-12
    # |dot| >= 1 - eps: plain linear blend of the coefficients (Eigen 3.3.7), no division by sin(theta) ~ 0
    a = wxyz(Rotation.from_rotvec([0.3, -0.2, 0.1]))
    out = np.zeros(4)
    lib.shim_quat_slerp(P(a), P(a.copy()), 0.37, P(out))
    np.testing.assert_allclose(out, a, atol=1e-15)
    lib.shim_quat_slerp(P(a), P(-a), 0.37, P(out))
    np.testing.assert_allclose(out, (1 - 0.37) * a - 0.37 * (-a), atol=1e-15)


def test_from_two_vectors_and_solve(lib):
    rng = np.random.default_rng(3)
    for _ in range(50):
        a, b = rng.normal(size=3), rng.normal(size=3)
        q = np.zeros(4)
        lib.shim_from_two_vectors(P(a), P(b), P(q))
        r = Rotation.from_quat([q[1], q[2], q[3], q[0]])
        np.testing.assert_allclose(r.apply(a / np.linalg.norm(a)), b / np.linalg.norm(b), atol=1e-12)
        assert abs(np.linalg.norm(q) - 1) < 1e-14
    m = rng.normal(size=(7, 7))
    spd = m @ m.T + 7 * np.eye(7)
    rhs, x = rng.normal(size=7), np.zeros(7)
    lib.shim_solve(P(spd), P(rhs), 7, P(x))
    np.testing.assert_allclose(x, np.linalg.solve(spd, rhs), atol=1e-12)
