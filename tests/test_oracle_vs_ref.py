"""Pins the oracle's restated pair kernels (oracle/oracle_impl.inc) against the
REFERENCE's own pair kernels: live against oracle/_ref (compiled from
/root/reference/src/c_rigid_obj.cpp:31-142 by oracle/build_ref.sh) where that .so exists,
and against tests/golden/pair_golden.npz (outputs of that same .so, committed) everywhere.
Bit-for-bit in float64 and float32."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden


def _oracle_blocks(orc, g, dt, wall):
    """Evaluate orc.pair_block on two-blob configurations reproducing the golden arguments
    (a = 1: blob i at (d, hi), blob j at (0, hj) -> rz + 2 z_j = hi + hj, h_j = hj)."""
    n = g["d"].shape[0]
    out = np.zeros((n, 9), dt)
    for k in range(n):
        d = g["d"][k].astype(dt)
        if wall:
            zi, zj = dt(g["hi"][k]), dt(g["hj"][k])
            r = np.array([d[0], d[1], zi, 0, 0, zj], dt)
        else:
            r = np.array([d[0], d[1], d[2], 0, 0, 0], dt)
        out[k] = orc.pair_block(r, 0, 1, 1.0, wall, dtype=dt).reshape(-1)
    return out


@pytest.mark.parametrize("sfx,dt", [("f64", np.float64), ("f32", np.float32)])
def test_rpy_pair_matches_reference_golden(orc, sfx, dt):
    g = load_golden("pair_golden")
    B = _oracle_blocks(orc, g, dt, wall=False)
    ref = g[f"rpy_{sfx}"]  # Mxx Mxy Mxz Myy Myz Mzz
    got = B[:, [0, 1, 2, 4, 5, 8]]
    assert np.array_equal(got, ref)
    assert np.array_equal(B[:, [3, 6, 7]], B[:, [1, 2, 5]])  # symmetric block (:437-439)


@pytest.mark.parametrize("sfx,dt", [("f64", np.float64), ("f32", np.float32)])
def test_wall_pair_matches_reference_golden(orc, sfx, dt):
    g = load_golden("pair_golden")
    n = g["d"].shape[0]
    # wall correction alone = (RPY + wall) - RPY is not bit-stable; rebuild the same sum the
    # reference loop forms instead: golden RPY block + golden wall correction, added in `dt`
    got = np.zeros((n, 9), dt)
    for k in range(n):
        d = g["d"][k].astype(dt)
        zi, zj = dt(g["hi"][k]), dt(g["hj"][k])
        dz = zi - zj
        r = np.array([d[0], d[1], zi, 0, 0, zj], dt)
        got[k] = orc.pair_block(r, 0, 1, 1.0, True, dtype=dt).reshape(-1)
        free = orc.pair_block(np.array([d[0], d[1], dz, 0, 0, 0], dt), 0, 1, 1.0, False, dtype=dt).reshape(-1)
        # the wall kernel sees (rz + 2 zj)/a computed from rz = zi - zj (:442)
        Rz = dt(dt(dz + dt(2) * zj) / dt(1.0))
        if Rz != dt(zi + zj):
            continue  # rounding made the golden argument differ; covered by the live test
        want = (free + g[f"wall_{sfx}"][k]).astype(dt)
        assert np.array_equal(got[k], want), k


@pytest.mark.parametrize("sfx,dt", [("f64", np.float64), ("f32", np.float32)])
def test_wall_self_matches_reference_golden(orc, sfx, dt):
    g = load_golden("pair_golden")
    for k in range(0, g["d"].shape[0], 7):
        zj = dt(g["hj"][k])
        r = np.array([0, 0, zj], dt)
        got = orc.pair_block(r, 0, 0, 1.0, True, dtype=dt).reshape(-1)
        want = np.zeros(9, dt)
        want[[0, 4, 8]] = dt(4.0) / dt(3.0)
        want = (want + g[f"wall_self_{sfx}"][k]).astype(dt)
        assert np.array_equal(got, want)


def test_below_wall_is_an_error(orc):
    g = load_golden("pair_golden")
    assert int(g["below_status"]) == 2  # the reference threw std::runtime_error
    r = np.array([0.0, 0.0, 1.0, 0.3, 0.0, -0.5])
    with pytest.raises(orc.OracleError):
        orc.pair_block(r, 0, 1, 1.0, True)
    with pytest.raises(orc.OracleError):
        orc.apply_M(np.ones(6), r, 1.0, 1.0, True)
    # without the wall the same configuration is fine
    orc.apply_M(np.ones(6), r, 1.0, 1.0, False)


@pytest.mark.parametrize("sfx,ct,dt", [("f64", ctypes.c_double, np.float64), ("f32", ctypes.c_float, np.float32)])
def test_live_against_compiled_reference(orc, sfx, ct, dt):
    R = orc.ref_pair_lib()
    if R is None:
        pytest.skip("oracle/_ref/libref_pair.so not present (needs /root/reference to build)")
    rng = np.random.default_rng(7)
    for _ in range(300):
        a = dt(rng.uniform(0.2, 1.5))
        ri = rng.uniform(-3, 3, 3).astype(dt)
        rj = rng.uniform(-3, 3, 3).astype(dt)
        ri[2], rj[2] = abs(ri[2]) + dt(0.05), abs(rj[2]) + dt(0.05)
        r = np.concatenate([ri, rj]).astype(dt)
        got = orc.pair_block(r, 0, 1, float(a), True, dtype=dt).reshape(-1)
        # drive the reference kernels exactly as rotne_prager_tensor does (:420-445)
        inv_a = dt(1.0) / a
        rx, ry, rz = ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2]
        m6 = np.zeros(6, dt)
        getattr(R, f"ref_rpy_pair_{sfx}")(ct(rx), ct(ry), ct(rz), m6.ctypes.data, 0, 1, ct(inv_a))
        m9 = np.array([m6[0], m6[1], m6[2], m6[1], m6[3], m6[4], m6[2], m6[4], m6[5]], dt)
        st = getattr(R, f"ref_wall_pair_{sfx}")(ct(dt(rx / a)), ct(dt(ry / a)), ct(dt(dt(rz + dt(2) * rj[2]) / a)),
                                                 m9.ctypes.data, 0, 1, ct(dt(rj[2] / a)))
        assert st == 0
        assert np.array_equal(got, m9)


# ---- the reference's own apply_M (assembly loop + B M B F), not just its pair kernels -------------
from conftest import CASE_NAMES, rel_err  # noqa: E402


@pytest.mark.parametrize("dt,sfx,tol", [(np.float64, "f64", 0.0), (np.float32, "f32", 0.0)])
@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_apply_M_equals_the_reference_members_golden(orc, name, dt, sfx, tol):
    """oracle.apply_M_dense (the restatement of rotne_prager_tensor + make_damp_mat + apply_M) against
    tests/golden/apply_M_ref_golden.npz = outputs of THOSE REFERENCE MEMBERS compiled from the
    reference source (oracle/build_ref.sh; Eigen replaced by oracle/eigen_shim.inc).  Bit for bit in
    both precisions: same pair kernels, same assembly order, same column-oriented GEMV.  The
    matrix-free oracle (long-double row sums) agrees to rounding."""
    g, ref = load_golden(name), load_golden("apply_M_ref_golden")
    a, eta, wall = float(g["a"]), float(g["eta"]), bool(g["wall"])
    want = ref[f"{name}/{sfx}"]
    got = orc.apply_M_dense(g["lam"].astype(dt), g["r"].astype(dt), a, eta, wall, dtype=dt)
    assert np.array_equal(got, want)
    mf = orc.apply_M(g["lam"].astype(dt).astype(np.float64), g["r"].astype(dt).astype(np.float64), a, eta, wall)
    assert rel_err(mf, want) < (1e-14 if dt == np.float64 else 1e-6)
    if dt == np.float64:
        assert rel_err(g["MF"], want) < 1e-14  # the fixture every GPU parity test compares with


@pytest.mark.parametrize("wall", [False, True])
def test_oracle_apply_M_equals_the_reference_members_on_a_ragged_cloud(orc, wall):
    ref = load_golden("apply_M_ref_golden")
    r, F, a, eta = ref["cloud/r"], ref["cloud/F"], float(ref["cloud/a"]), float(ref["cloud/eta"])
    assert np.array_equal(orc.apply_M_dense(F, r, a, eta, wall), ref[f"cloud/wall{int(wall)}/f64"])
    assert rel_err(orc.apply_M(F, r, a, eta, wall), ref[f"cloud/wall{int(wall)}/f64"]) < 1e-14
    f32 = orc.apply_M_dense(F.astype(np.float32), r.astype(np.float32), a, eta, wall, dtype=np.float32)
    assert np.array_equal(f32, ref[f"cloud/wall{int(wall)}/f32"])


def test_reference_members_live_when_the_reference_tree_is_here(orc):
    """where oracle/_ref/libref_apply_M.so exists (this container; it travels to the GPU box too):
    the live library reproduces its committed golden outputs, and throws below the wall."""
    if orc.ref_apply_M_lib() is None:
        pytest.skip("oracle/_ref/libref_apply_M.so not built (no /root/reference here)")
    g, ref = load_golden("case_touch_wall"), load_golden("apply_M_ref_golden")
    u = orc.ref_apply_M(g["lam"], g["r"], float(g["a"]), float(g["eta"]), True)
    assert np.array_equal(u, ref["case_touch_wall/f64"])
    r = g["r"].copy()
    r[0, 2] = -0.1
    with pytest.raises(orc.OracleError):
        orc.ref_apply_M(g["lam"], r, float(g["a"]), float(g["eta"]), True)


# ---- the reference's own state / placement / K / integrator members ----------------------------------
@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_rigid_members_equal_the_reference_members_golden(orc, name):
    """oracle.py's numpy restatement of setConfig (normalisation), multi_body_pos, K_x_U, KT_x_Lam,
    Kinv_x_V, KTinv_x_F and evolve_X_Q against tests/golden/members_ref_golden.npz = outputs of those
    REFERENCE MEMBERS compiled from the reference source (oracle.RefBody).  To rounding (the sums are a
    handful of terms; the reference's sparse products and the oracle's closed forms order them
    differently); the float build of the reference agrees at float accuracy."""
    g, ref = load_golden(name), load_golden("members_ref_golden")
    rcfg = orc.remove_mean(g["cfg"])
    n_blb = rcfg.shape[0]
    Qn = orc.normalize_quats(g["Q"])
    r = orc.blob_positions(g["X"], Qn, rcfg)
    got = {"Qn": Qn, "r": r, "KU": orc.K_dot(g["U"], r, g["X"], n_blb), "KTlam": orc.KT_dot(g["lam"], r, g["X"], n_blb),
           "Kinv_lam": orc.Kinv_dense(r, g["X"], Qn, rcfg) @ g["lam"], "KinvT_U": orc.Kinv_dense(r, g["X"], Qn, rcfg).T @ g["U"]}
    Xe, Qe = orc.evolve(g["X"], Qn, g["U"], float(g["dt"]))
    got["X_evolved"], got["Q_evolved"] = Xe, Qe
    got["KU_evolved"] = orc.K_dot(g["U"], orc.blob_positions(Xe, Qe, rcfg), Xe, n_blb)
    for key, val in got.items():
        assert rel_err(val, ref[f"{name}/f64/{key}"]) < 2e-15, key
        assert rel_err(val, ref[f"{name}/f32/{key}"]) < 5e-6, key
    # and the committed case fixtures the GPU tests use ARE these values
    for key in ("Qn", "r", "KU", "KTlam", "Kinv_lam", "KinvT_U", "X_evolved", "Q_evolved"):
        assert rel_err(g[key], ref[f"{name}/f64/{key}"]) < 2e-15, key
    # both preconditioners: apply_PC of the reference (Eigen's inverse() / LLT through the shim) vs the
    # oracle's PC class, on the cases where the reference's own LLT does not break down
    if "pc_diag" in g:
        for blk, key, tol in ((False, "pc_diag", 1e-14), (True, "pc_block", 1e-12)):
            mine = orc.PC(g["X"], Qn, rcfg, float(g["a"]), float(g["eta"]), bool(g["wall"]), blk).apply(g["vec"])
            assert rel_err(mine, ref[f"{name}/f64/{key}"]) < tol, key
            assert rel_err(g[key], ref[f"{name}/f64/{key}"]) < tol, key
            assert rel_err(mine, ref[f"{name}/f32/{key}"]) < 5e-3, key
    else:  # blobs inside the wall-overlap layer: the reference's LLT of K^T Mt^-1 K yields NaNs
        assert not np.isfinite(ref[f"{name}/f64/pc_diag"]).all()


def test_reference_member_library_live(orc):
    if orc.ref_apply_M_lib() is None:
        pytest.skip("oracle/_ref/libref_members.so not built (no /root/reference here)")
    g, ref = load_golden("case_touch_wall"), load_golden("members_ref_golden")
    rb = orc.RefBody(g["cfg"], g["X"], g["Q"], float(g["a"]), float(g["eta"]), float(g["dt"]), wall_PC=True)
    assert np.array_equal(rb.positions(), ref["case_touch_wall/f64/r"])
    assert np.array_equal(rb.K_dot(g["U"]), ref["case_touch_wall/f64/KU"])
    assert np.array_equal(rb.KT_dot(g["lam"]), ref["case_touch_wall/f64/KTlam"])
    # the reference's apply_M accepts more blobs than the bodies hold (tests/test_interface.py:171-177)
    r = np.concatenate([rb.positions().reshape(-1), [7.0, -3.0, 2.0]])
    F = np.concatenate([g["lam"], [0.3, -0.2, 0.9]])
    assert rel_err(rb.apply_M(F, r), orc.apply_M(F, r, float(g["a"]), float(g["eta"]), True)) < 1e-14


@pytest.mark.parametrize("seed", range(12))
def test_oracle_equals_the_reference_members_on_random_suspensions(orc, seed):
    """Beyond the five committed cases: seeded random suspensions (1-4 bodies, random blob clouds as
    body shapes, random radius / viscosity / un-normalised quaternions, with and without the wall)
    through the live reference-member library and through the oracle.  Skipped where
    oracle/_ref/libref_members.so does not exist."""
    if orc.ref_apply_M_lib() is None:
        pytest.skip("oracle/_ref/libref_members.so not built (no /root/reference here)")
    rng = np.random.default_rng(1000 + seed)
    nb, n_blb, wall = int(rng.integers(1, 5)), int(rng.integers(3, 9)), bool(seed % 2)
    cfg = rng.uniform(-0.8, 0.8, (n_blb, 3))
    a, eta, dt = float(rng.uniform(0.1, 0.4)), float(rng.uniform(0.5, 2.0)), 0.01
    X = rng.uniform(-3, 3, (nb, 3))
    X[:, 2] = rng.uniform(2.0, 4.0, nb)  # every blob above z = a
    Q = rng.standard_normal((nb, 4)) * rng.uniform(0.3, 3.0, (nb, 1))
    rb = orc.RefBody(cfg, X, Q, a, eta, dt, wall_PC=wall)
    rcfg, Qn = orc.remove_mean(cfg), orc.normalize_quats(Q)
    r = orc.blob_positions(X, Qn, rcfg)
    lam, U = rng.standard_normal(3 * nb * n_blb), rng.standard_normal(6 * nb)
    vec = rng.standard_normal(3 * nb * n_blb + 6 * nb)
    assert rel_err(rb.get_config()[1], Qn) < 1e-15 and rel_err(rb.positions(), r) < 1e-15
    assert rel_err(orc.K_dot(U, r, X, n_blb), rb.K_dot(U)) < 5e-15
    assert rel_err(orc.KT_dot(lam, r, X, n_blb), rb.KT_dot(lam)) < 5e-15
    Kinv = orc.Kinv_dense(r, X, Qn, rcfg)
    assert rel_err(Kinv @ lam, rb.Kinv_dot(lam)) < 1e-13 and rel_err(Kinv.T @ U, rb.KTinv_dot(U)) < 1e-13
    assert np.array_equal(orc.apply_M_dense(lam, r, a, eta, wall), rb.apply_M(lam, r))
    assert rel_err(orc.apply_M(lam, r, a, eta, wall), rb.apply_M(lam, r)) < 1e-14
    saddle = np.concatenate([rb.apply_M(vec[:lam.size], rb.positions()) - rb.K_dot(vec[lam.size:]), rb.KT_dot(vec[:lam.size])])
    assert rel_err(orc.apply_saddle(vec, X, Qn, rcfg, a, eta, wall), saddle) < 1e-14  # Rigid.py:73-80 composition
    for blk in (False, True):
        ref_pc = orc.RefBody(cfg, X, Q, a, eta, dt, wall_PC=wall, block_PC=blk).apply_PC(vec)
        if np.isfinite(ref_pc).all():
            assert rel_err(orc.PC(X, Qn, rcfg, a, eta, wall, blk).apply(vec), ref_pc) < 1e-10
    rb.evolve(U)
    Xe, Qe = orc.evolve(X, Qn, U, dt)
    assert rel_err(rb.get_config()[0], Xe) < 1e-15 and rel_err(rb.get_config()[1], Qe) < 1e-15


@pytest.mark.parametrize("name", ["case_touch_wall", "case_touch_free", "case_overlap_wall"])
def test_rfd_and_brownian_increment_against_the_reference_members(orc, name):
    """The stochastic pieces of the BD step against the reference's own M_RFD (:769-796) and M_half_W
    (:661-675), compiled from its source with rand_vector returning an injected noise vector (the
    reference seeds it from the wall clock): the oracle's random finite difference is the reference's
    for the same noise, and the reference's Brownian increment chol(B M B) W has the covariance the
    drop-in's increments have (a different square root of the same matrix)."""
    if orc.ref_apply_M_lib() is None:
        pytest.skip("oracle/_ref/libref_members.so not built (no /root/reference here)")
    g = load_golden(name)
    a, eta, wall, dt = float(g["a"]), float(g["eta"]), bool(g["wall"]), float(g["dt"])
    rb = orc.RefBody(g["cfg"], g["X"], g["Q"], a, eta, dt, wall_PC=wall)
    rcfg = orc.remove_mean(g["cfg"])
    W = np.random.default_rng(3).standard_normal(g["r"].size)
    mine = orc.rfd_M(g["X"], g["Qn"], rcfg, a, eta, wall, W, delta=1.0e-4)
    # a difference quotient with delta = 1e-4 amplifies rounding by 1e4: agreement to ~1e-10
    assert rel_err(mine, rb.M_RFD(W)) < 1e-8
    n = g["r"].size
    if n <= 200:
        A = np.asarray(orc.dense_mobility(g["r"], a, eta, wall))
        B = orc.damp_diag(g["r"], a)
        A = B[:, None] * A * B[None, :]
        L = np.stack([rb.M_half_W(e) for e in np.eye(n)], axis=1)
        assert np.allclose(L, np.tril(L)) and np.linalg.norm(L @ L.T - A) / np.linalg.norm(A) < 1e-13
        fac = orc.noise_factors(g["r"], g["Qn"], rcfg, a, eta, wall)
        S = np.stack([orc.noise_block_cholesky(fac, A, e) for e in np.eye(n)], axis=1)
        assert np.linalg.norm(S @ S.T - L @ L.T) / np.linalg.norm(A) < 1e-9  # same covariance, different roots
        assert np.linalg.norm(S - L) / np.linalg.norm(L) > 1e-2


@pytest.mark.parametrize("name", ["case_touch_wall", "case_touch_free"])
def test_bd_right_hand_side_against_the_reference_RHS_and_Midpoint(orc, name):
    """The right-hand side of the BD step's saddle solve against the reference's own RHS_and_Midpoint
    (:917-976) compiled from its source, fed the same three noise vectors in the order it draws them
    and the reference's own noise route (Cholesky factor): slip row  slip - kBT M_RFD - sqrt(kBT/dt)
    (M^{1/2}W1 - M^{1/2}W2)  identical; force row: the reference returns -F for its PC's
    [M -K; -K^T 0] convention, the drop-in solves apply_saddle = [M -K; +K^T 0] with +F
     -- the same equations."""
    if orc.ref_apply_M_lib() is None:
        pytest.skip("oracle/_ref/libref_members.so not built (no /root/reference here)")
    g = dict(load_golden(name))
    a, eta, wall, dt = float(g["a"]), float(g["eta"]), bool(g["wall"]), float(g["dt"])
    rcfg = orc.remove_mean(g["cfg"])
    if not wall:
        # the reference's M_half_W multiplies by the wall damping B = min(1, z/a) even without the wall
        # (:668-669; the drop-in applies B with the wall only): lift the free-space
        # case so that every blob has z >= a and B = I, which is what this test is not about
        g["X"] = g["X"] + np.array([0.0, 0.0, 5.0])
        g["r"] = orc.blob_positions(g["X"], g["Qn"], rcfg)
        assert g["r"][:, 2].min() >= a
    rb = orc.RefBody(g["cfg"], g["X"], g["Q"], a, eta, dt, wall_PC=wall)
    rng = np.random.default_rng(31)
    nb, n3 = g["X"].shape[0], g["r"].size
    F, slip = rng.standard_normal(6 * nb), 0.1 * rng.standard_normal(n3)
    W = [rng.standard_normal(n3) for _ in range(3)]
    kBT = 0.004
    ref_rhs = rb.RHS_and_Midpoint(slip, F, *W, kBT)
    rhs, Xm, Qm = orc.bd_step(g["X"], g["Qn"], rcfg, a, eta, dt, kBT, wall, F, slip, *W, noise="cholesky", return_rhs=True)
    assert rel_err(rhs[:n3], ref_rhs[:n3]) < 1e-9
    assert np.array_equal(ref_rhs[n3:], -F) and np.array_equal(rhs[n3:], F)
    # the step moved to the midpoint along K^-1 (2 sqrt(kBT/dt) M^{1/2} W1) dt/2  (:954-958)
    Lc = np.linalg.cholesky(np.asarray(orc.dense_mobility(g["r"], a, eta, wall)) * np.outer(orc.damp_diag(g["r"], a), orc.damp_diag(g["r"], a)))
    uom = orc.Kinv_apply(2.0 * np.sqrt(kBT / dt) * (Lc @ W[0]), g["r"], g["X"], g["Qn"], rcfg)
    Xw, Qw = orc.update_X_Q(g["X"], g["Qn"], 0.5 * dt * uom)
    assert rel_err(Xm, Xw) < 1e-14 and rel_err(Qm, Qw) < 1e-14
    # the single-increment branch (split_rand = false, :949-953): rand_vector is drawn for W1, then for M_RFD
    rb.set_split_rand(False)
    ref_rhs1 = rb.RHS_and_Midpoint(slip, F, W[0], W[2], W[2], kBT)  # injected order: W1, Wr (third unused)
    rhs1, _, _ = orc.bd_step(g["X"], g["Qn"], rcfg, a, eta, dt, kBT, wall, F, slip, W[0], None, W[2], noise="cholesky",
                             return_rhs=True, split_rand=False)
    assert rel_err(rhs1[:n3], ref_rhs1[:n3]) < 1e-9
    assert rel_err(rhs1[:n3], rhs[:n3]) > 1e-3  # a different right-hand side than the two-increment scheme
