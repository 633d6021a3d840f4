"""The two-right-hand-side symmetric product (rpy_matvec_sym2_kernel, rbl_apply_M2) and the
paired Lanczos / BD step built on it, against the oracle and against the single-vector path."""
import ctypes

import numpy as np
import pytest

from conftest import CASE_NAMES, TOL, load_golden, rel_err, check
from test_gpu_matvec import _dtype, _oracle_for, _random_cloud, _solver

pytestmark = pytest.mark.gpu
PRECISIONS = ["double", "single"]


def _apply_M2(ctx, F1, F2, r):
    F1, F2, r = ctx._arr(F1), ctx._arr(F2), ctx._arr(r)
    o1, o2 = np.empty_like(F1), np.empty_like(F2)
    ctx.call("rbl_apply_M2", F1.ctypes.data, F2.ctypes.data, r.ctypes.data, F1.size // 3, o1.ctypes.data, o2.ctypes.data)
    return o1, o2


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_apply_M2_matches_oracle_on_golden_cases(orc, name, precision):
    g = load_golden(name)
    cb = _solver(g, precision)
    F2 = np.random.default_rng(5).standard_normal(g["lam"].size)
    o1, o2 = cb.apply_M2(g["lam"], F2, g["r"])
    assert o1.dtype == _dtype(precision)
    a, eta, wall = float(g["a"]), float(g["eta"]), bool(g["wall"])
    check(rel_err(o1, _oracle_for(orc, g["lam"], g["r"], a, eta, wall, precision)), TOL[precision])
    check(rel_err(o2, _oracle_for(orc, F2, g["r"], a, eta, wall, precision)), TOL[precision])


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("wall", [False, True])
@pytest.mark.parametrize("n", [1, 2, 33, 256, 257, 1025, 4099])
def test_apply_M2_ragged_sizes_and_variants(orc, n, wall, precision):
    from rigid_body_light_b200._lib import Context

    a, eta = 0.11, 0.9
    r, F1 = _random_cloud(n, wall, seed=n)
    F2 = np.random.default_rng(n + 1).standard_normal(3 * n)
    ctx = Context(precision)
    ctx.set_parameters(a, 0.01, 1.0, eta, np.zeros((1, 3)))
    ctx.set_flags(0, wall)
    w1 = _oracle_for(orc, F1, r, a, eta, wall, precision)
    w2 = _oracle_for(orc, F2, r, a, eta, wall, precision)
    nv = ctx.L.rbl_num_sym2_variants(ctx.h)
    assert nv >= 2
    for v in [-1] + list(range(nv)):
        ctx.call("rbl_set_sym2_variant", v)
        o1, o2 = _apply_M2(ctx, F1, F2, r)
        check(rel_err(o1, w1), TOL[precision])
        check(rel_err(o2, w2), TOL[precision])
    ctx.close()


@pytest.mark.parametrize("precision", PRECISIONS)
def test_apply_M2_equals_two_single_products_on_a_suspension(orc, precision):
    """27 touching spheres of shell_N_162 above the wall (4374 blobs: several target tiles, far
    and near tile pairs): the two-vector pass equals two single-vector passes to rounding."""
    from Rigid import RigidBody
    from rigid_body_light_b200.shells import sphere_suspension

    s = sphere_suspension(27, 162, True)
    cb = RigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, precision=precision)
    r = cb.get_blob_positions()
    rng = np.random.default_rng(3)
    F1, F2 = rng.standard_normal(r.size), rng.standard_normal(r.size)
    o1, o2 = cb.apply_M2(F1, F2, r)
    tol = 1e-13 if precision == "double" else 5e-6
    check(rel_err(o1, cb.apply_M(F1, r)), tol)
    check(rel_err(o2, cb.apply_M(F2, r)), tol)
    rows = np.random.default_rng(4).choice(r.shape[0], 64, replace=False)
    want = _oracle_for(orc, F2, r, s["a"], 1.0, True, precision, rows=rows)
    check(rel_err(o2.reshape(-1, 3)[rows], want), TOL[precision])


def test_blob_below_wall_raises_in_the_two_rhs_product():
    from rigid_body_light_b200._lib import Context, RblError

    ctx = Context("double")
    ctx.set_parameters(0.1, 0.01, 1.0, 1.0, np.zeros((1, 3)))
    ctx.set_flags(0, 1)
    r = np.array([[0, 0, 0.5], [1, 0, -0.01]])
    with pytest.raises(RblError) as e:
        _apply_M2(ctx, np.ones(6), np.ones(6), r)
    assert e.value.status == 3
    ctx.close()


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", ["case_touch_wall", "case_overlap_free"])
def test_paired_lanczos_equals_two_single_runs(name, precision):
    g = load_golden(name)
    cb = _solver(g, precision)
    rng = np.random.default_rng(9)
    n = g["r"].size
    W1, W2 = rng.standard_normal(n), 3.0 * rng.standard_normal(n)
    tol = 1e-10 if precision == "double" else 1e-5
    y1, y2, k1, k2 = cb.brownian_sqrt_pair(W1, W2, tol=tol, max_iter=150)
    s1, j1 = cb.brownian_sqrt(W1, tol=tol, max_iter=150)
    s2, j2 = cb.brownian_sqrt(W2, tol=tol, max_iter=150)
    lim = 1e-13 if precision == "double" else 2e-6  # observed 4.6e-15 / 1.4e-7
    check(rel_err(y1, s1), lim)
    check(rel_err(y2, s2), lim)
    assert abs(k1 - j1) <= 1 and abs(k2 - j2) <= 1
    # one vector zero: that recurrence is skipped, the other is unaffected
    z1, z2, m1, m2 = cb.brownian_sqrt_pair(np.zeros(n), W2, tol=tol, max_iter=150)
    assert m1 == 0
    assert not z1.any()
    check(rel_err(z2, s2), lim)


def test_bd_step_paired_and_unpaired_lanczos_agree():
    g = load_golden("case_touch_wall")
    nb, n3 = g["X"].shape[0], g["r"].size
    rng = np.random.default_rng(31)
    F = rng.standard_normal(6 * nb)
    noise = tuple(rng.standard_normal(n3) for _ in range(3))
    out = []
    for pairing in (1, 0):
        cb = _solver(g, "double", block=True)
        from rigid_body_light_b200._lib import load

        lib = load()
        assert lib.rbl_set_lanczos_pairing(ctypes.c_void_p(cb.cb.handle()), pairing) == 0
        U, it, rr = cb.bd_step(F, kBT=0.004, noise=noise, tol=1e-11, restart=100, max_iter=400, lanczos_tol=1e-12,
                               lanczos_max_iter=200)
        out.append((U, cb.get_config()))
    check(rel_err(out[0][0], out[1][0]), 1e-13)  # observed 4e-15
    check(rel_err(out[0][1][0], out[1][1][0]), 1e-14)


def test_full_size_krylov_properties_config3():
    """BASELINE.json configs[2] size (4096 spheres x shell_N_42 = 172 032 blobs, wall), float:
    size-independent properties of the Krylov drivers where the dense oracle cannot go --
    the square root applied twice is the mobility product, the paired run equals itself with the
    vectors swapped, and the GMRES solution satisfies the saddle system (true residual through an
    independent operator application)."""
    from Rigid import RigidBody
    from rigid_body_light_b200.shells import sphere_suspension

    s = sphere_suspension(4096, 42, True)
    cb = RigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, block_PC=True, precision="single")
    n3, n6 = 3 * 4096 * 42, 6 * 4096
    rng = np.random.default_rng(17)
    W1, W2 = rng.standard_normal(n3), rng.standard_normal(n3)
    y1, y2, k1, k2 = cb.brownian_sqrt_pair(W1, W2, tol=1e-5, max_iter=80)
    z1, z2, _, _ = cb.brownian_sqrt_pair(y1, y2, tol=1e-5, max_iter=80)
    r = cb.get_blob_positions()
    m1, m2 = cb.apply_M2(W1, W2, r)
    # Krylov convergence, not parity: each square root stops at a relative change of 1e-5; observed 4.4e-5
    check(rel_err(z1, m1), 4e-4, "S(S W) vs M W with the Lanczos stopping tolerance 1e-5, float")
    check(rel_err(z2, m2), 4e-4, "S(S W) vs M W with the Lanczos stopping tolerance 1e-5, float")
    s2, s1, j2, j1 = cb.brownian_sqrt_pair(W2, W1, tol=1e-5, max_iter=80)
    assert (j1, j2) == (k1, k2)
    check(rel_err(s1, y1), 5e-6)  # observed 3.5e-7
    check(rel_err(s2, y2), 5e-6)
    rhs = np.concatenate([np.zeros(n3), rng.standard_normal(n6)]).astype(np.float32)
    x, it, rr = cb.gmres(rhs, tol=1e-4, restart=60, max_iter=120)
    assert rr <= 1e-4 and it < 60
    check(rel_err(cb.apply_saddle(x), rhs), 2e-4, "true residual of the float GMRES solution (requested 1e-4)")  # observed 6.2e-5
