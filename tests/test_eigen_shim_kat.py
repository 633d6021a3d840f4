"""Known-answer test of oracle/eigen_shim.inc, the from-scratch subset of Eigen through which the reference's
own member functions are compiled into oracle/_ref/libref_members.so (Eigen3 is not installed here).  The
driver tests/host/shim_kat.cpp runs every operation those members use on fixed inputs; this file recomputes
them with numpy / scipy under Eigen's DOCUMENTED conventions -- column-major storage, Quaternion
coefficients (x, y, z, w) with the Hamilton product and the standard rotation-matrix formula, duplicate
triplets summed by setFromTriplets, normalize() leaving a zero quaternion untouched, LLT = lower Cholesky
factor.  CPU only (g++); pins the shim's semantics independently of the reference members that use it."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def kat(tmp_path_factory):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("shim") / "shim_kat")
    src = os.path.join(ROOT, "tests", "host", "shim_kat.cpp")
    subprocess.run([cxx, "-O1", "-std=c++17", "-o", exe, src], check=True, capture_output=True, text=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return json.loads(out)


def _entry(n, m):
    i, j = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
    return np.sin(1.0 + 0.7 * i + 1.3 * j) + 2.0 * (i == j)


def _mat(d):
    """dense matrix from the raw COLUMN-MAJOR dump"""
    return np.asarray(d["raw"], dtype=np.float64).reshape(d["cols"], d["rows"]).T


def test_storage_is_column_major(kat):
    # M(i, j) = 10 i + j in a 2 x 3 matrix lies in memory column by column (Eigen's default order)
    assert kat["storage_2x3"]["raw"] == [0, 10, 1, 11, 2, 12]


def test_dense_algebra(kat):
    A3, A6, B = _entry(3, 3), _entry(6, 6), _entry(6, 3)
    np.testing.assert_allclose(_mat(kat["inv3"]), np.linalg.inv(A3), rtol=0, atol=1e-14)
    np.testing.assert_allclose(_mat(kat["inv6"]), np.linalg.inv(A6), rtol=0, atol=1e-13)
    assert abs(kat["det3"] - np.linalg.det(A3)) < 1e-14
    np.testing.assert_array_equal(_mat(kat["transpose63"]), B.T)
    np.testing.assert_allclose(_mat(kat["prod66_63"]), A6 @ B, rtol=0, atol=1e-14)
    np.testing.assert_allclose(_mat(kat["sum"]), A6 + A6.T, rtol=0, atol=1e-15)
    np.testing.assert_allclose(_mat(kat["diff"]), A6 - A6.T, rtol=0, atol=1e-15)
    np.testing.assert_allclose(_mat(kat["scaled"]), 2.5 * A3, rtol=0, atol=1e-15)
    np.testing.assert_allclose(_mat(kat["colmean"]).ravel(), B.mean(axis=0), rtol=0, atol=1e-15)
    assert abs(kat["norm"] - np.linalg.norm(B)) < 1e-14 and abs(kat["sqnorm"] - (B ** 2).sum()) < 1e-13
    np.testing.assert_array_equal(_mat(kat["block"]), A6[1:4, 2:4])
    d = 0.5 + 0.1 * np.arange(6)
    np.testing.assert_allclose(_mat(kat["diag_left"]), np.diag(d) @ A6, rtol=0, atol=1e-15)   # rows scaled
    np.testing.assert_allclose(_mat(kat["diag_right"]), A6 @ np.diag(d), rtol=0, atol=1e-15)  # columns scaled


def test_llt_is_the_lower_cholesky_factor(kat):
    A6 = _entry(6, 6)
    P = A6 @ A6.T + 6.0 * np.eye(6)
    L = _mat(kat["chol_L"])
    np.testing.assert_allclose(L, np.linalg.cholesky(P), rtol=0, atol=1e-13)
    assert np.all(np.triu(L, 1) == 0)
    rhs = np.cos(0.3 * np.arange(6))
    np.testing.assert_allclose(_mat(kat["chol_solve"]).ravel(), np.linalg.solve(P, rhs), rtol=0, atol=1e-14)


def test_quaternion_conventions(kat):
    from scipy.spatial.transform import Rotation

    q = np.array([-0.5, 0.2, 0.7, 0.3])  # (x, y, z, w): Eigen's coefficient order, scipy's too
    q /= np.linalg.norm(q)
    p = np.array([0.1, 0.9, -0.3, 0.4])
    p /= np.linalg.norm(p)
    np.testing.assert_allclose(_mat(kat["quat_rot"]), Rotation.from_quat(q).as_matrix(), rtol=0, atol=1e-15)
    # Hamilton product: rotation by p first, then by q  (scipy composes the same way: (Rq * Rp) applies Rp first)
    want = (Rotation.from_quat(q) * Rotation.from_quat(p)).as_quat()
    got = np.array(kat["quat_prod_xyzw"])
    assert min(np.abs(got - want).max(), np.abs(got + want).max()) < 1e-15
    assert kat["quat_zero_xyzw"] == [0, 0, 0, 0]      # normalize() leaves the zero quaternion untouched
    assert kat["quat_identity_xyzw"] == [0, 0, 0, 1]


def test_triplets_and_stacking(kat):
    T = np.zeros((3, 3))
    for i, j, v in [(0, 0, 1.0), (1, 2, 2.0), (1, 2, 0.5), (2, 1, -3.0), (0, 0, 0.25)]:
        T[i, j] += v  # duplicates are summed
    np.testing.assert_array_equal(_mat(kat["triplets"]), T)
    np.testing.assert_allclose(_mat(kat["sparse_prod"]), T @ T.T, rtol=0, atol=1e-15)
    B, A3 = _entry(6, 3), _entry(3, 3)
    np.testing.assert_array_equal(np.asarray(kat["stacked"]["raw"]), np.concatenate([B[:, 0], A3[:, 1]]))
