"""The CUDA kernel's per-pair arithmetic (rigid_body_light_b200/csrc/rbl_pair.cuh: the
division-free, branch-free reformulation of c_rigid_obj.cpp:31-142) compiled for the HOST
and summed naively, against the oracle.  Checks the algebra of the kernel without a GPU;
it is not a CPU fallback (tests/host/ is never imported by the package)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import CASE_NAMES, ROOT, TOL, load_golden, rel_err


@pytest.fixture(scope="module")
def host_lib():
    src = os.path.join(ROOT, "tests", "host", "pair_host.cpp")
    so = os.path.join(ROOT, "tests", "host", "libpair_host.so")
    hdr = os.path.join(ROOT, "rigid_body_light_b200", "csrc", "rbl_pair.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src], check=True)
    return ctypes.CDLL(so)


def _host_matvec(L, F, r, a, eta, wall, near, dt, sym=False):
    F = np.ascontiguousarray(F, dt)
    r = np.ascontiguousarray(r, dt).reshape(-1)
    U = np.empty_like(F)
    name = "pair_matvec_sym_host_" if sym else "pair_matvec_host_"
    fn = getattr(L, name + ("f64" if dt == np.float64 else "f32"))
    fn(ctypes.c_void_p(F.ctypes.data), ctypes.c_void_p(r.ctypes.data), ctypes.c_int(r.size // 3), ctypes.c_double(a),
       ctypes.c_double(eta), ctypes.c_int(wall), ctypes.c_int(near), ctypes.c_void_p(U.ctypes.data))
    return U


@pytest.mark.parametrize("name", CASE_NAMES)
def test_kernel_pair_math_double(host_lib, name):
    g = load_golden(name)
    u = _host_matvec(host_lib, g["lam"], g["r"], float(g["a"]), float(g["eta"]), int(g["wall"]), 1, np.float64)
    assert rel_err(u, g["MF"]) < TOL["double"]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_kernel_pair_math_float(host_lib, orc, name):
    g = load_golden(name)
    a, eta, wall = float(g["a"]), float(g["eta"]), bool(g["wall"])
    # same inputs for both sides: the float32-representable positions and forces
    r32 = g["r"].astype(np.float32)
    f32 = g["lam"].astype(np.float32)
    want = orc.apply_M(f32.astype(np.float64), r32.astype(np.float64), a, eta, wall)
    u = _host_matvec(host_lib, f32, r32, a, eta, int(wall), 1, np.float32)
    assert rel_err(u, want) < TOL["single"]


def test_far_only_path_equals_general_path_when_separated(host_lib):
    """NEAR=false (overlap branch compiled out of the loop) is what the kernel runs on tile
    pairs whose bounding boxes are >= 2a apart; there it must equal the general path."""
    rng = np.random.default_rng(0)
    a = 0.2
    src = rng.uniform(0.3, 2.0, (40, 3))
    tgt = rng.uniform(0.3, 2.0, (25, 3)) + [2.0 + 2 * a, 0, 0]  # box gap >= 2a
    F = rng.standard_normal(src.size)
    for wall in (0, 1):
        out = []
        for near in (1, 0):
            U = np.empty(tgt.size)
            host_lib.pair_cross_host_f64(ctypes.c_void_p(tgt.ctypes.data), ctypes.c_int(25), ctypes.c_void_p(F.ctypes.data),
                                         ctypes.c_void_p(src.ctypes.data), ctypes.c_int(40), ctypes.c_double(a),
                                         ctypes.c_double(1.0), ctypes.c_int(wall), ctypes.c_int(near),
                                         ctypes.c_void_p(U.ctypes.data))
            out.append(U)
        assert np.array_equal(out[0], out[1])
    # ... and it is NOT interchangeable when blobs overlap
    tgt2 = src[:25] + 0.5 * a
    out = []
    for near in (1, 0):
        U = np.empty(tgt2.size)
        host_lib.pair_cross_host_f64(ctypes.c_void_p(tgt2.ctypes.data), ctypes.c_int(25), ctypes.c_void_p(F.ctypes.data),
                                     ctypes.c_void_p(src.ctypes.data), ctypes.c_int(40), ctypes.c_double(a),
                                     ctypes.c_double(1.0), ctypes.c_int(0), ctypes.c_int(near), ctypes.c_void_p(U.ctypes.data))
        out.append(U)
    assert rel_err(out[1], out[0]) > 1e-3


@pytest.mark.parametrize("name", CASE_NAMES)
def test_symmetric_pair_evaluation_double(host_lib, name):
    """pair_sym: one evaluation per unordered pair applied in both directions (the arithmetic
    of the symmetric CUDA kernel) reproduces the full ordered product."""
    g = load_golden(name)
    u = _host_matvec(host_lib, g["lam"], g["r"], float(g["a"]), float(g["eta"]), int(g["wall"]), 1, np.float64, sym=True)
    assert rel_err(u, g["MF"]) < TOL["double"]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_symmetric_pair_evaluation_float(host_lib, orc, name):
    g = load_golden(name)
    a, eta, wall = float(g["a"]), float(g["eta"]), bool(g["wall"])
    r32, f32 = g["r"].astype(np.float32), g["lam"].astype(np.float32)
    want = orc.apply_M(f32.astype(np.float64), r32.astype(np.float64), a, eta, wall)
    u = _host_matvec(host_lib, f32, r32, a, eta, int(wall), 1, np.float32, sym=True)
    assert rel_err(u, want) < TOL["single"]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name", CASE_NAMES)
def test_two_rhs_symmetric_pair_evaluation(host_lib, orc, name, dt):
    """pair_symR<2>: the geometry of an unordered pair evaluated once and applied to two force
    vectors in both directions (the two-right-hand-side kernel behind the paired Lanczos of the
    BD step) gives the two products."""
    g = load_golden(name)
    a, eta, wall = float(g["a"]), float(g["eta"]), bool(g["wall"])
    r = np.ascontiguousarray(g["r"].astype(dt)).reshape(-1)
    F1 = np.ascontiguousarray(g["lam"].astype(dt)).reshape(-1)
    F2 = np.ascontiguousarray(np.random.default_rng(7).standard_normal(F1.size).astype(dt))
    U1, U2 = np.empty_like(F1), np.empty_like(F2)
    fn = getattr(host_lib, "pair_matvec_sym2_host_" + ("f64" if dt == np.float64 else "f32"))
    p = lambda v: ctypes.c_void_p(v.ctypes.data)  # noqa: E731
    fn(p(F1), p(F2), p(r), ctypes.c_int(r.size // 3), ctypes.c_double(a), ctypes.c_double(eta), ctypes.c_int(wall),
       ctypes.c_int(1), p(U1), p(U2))
    tol = TOL["double"] if dt == np.float64 else TOL["single"]
    for F, U in ((F1, U1), (F2, U2)):
        want = orc.apply_M(F.astype(np.float64), r.astype(np.float64), a, eta, wall)
        assert rel_err(U, want) < tol
