"""Independent checks that defend the oracle where the reference pins nothing
(SURVEY.md section 8c): symmetry / positive definiteness of M, the analytic single-sphere
wall mobility, the Oseen far field, continuity at r = 2a, adjointness of K and K^T, the
preconditioner's exactness and sign convention (test_PC, c_rigid_obj.cpp:569-587), blob
placement against scipy exactly like /root/reference/tests/test_interface.py:55-73."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from conftest import CASE_NAMES, load_golden, rel_err


def _suspension(orc, nb=4, shell=12, wall=True, a=None, zshift=0.0):
    from rigid_body_light_b200.shells import sphere_suspension

    s = sphere_suspension(nb, shell, wall)
    s["X"][:, 2] += zshift
    ref = orc.remove_mean(s["cfg"])
    r = orc.blob_positions(s["X"], s["Q"], ref)
    return s, ref, r, (a if a is not None else s["a"])


@pytest.mark.parametrize("wall", [False, True])
def test_mobility_symmetric_positive_definite(orc, wall):
    s, ref, r, a = _suspension(orc, 4, 12, wall)
    M = np.asarray(orc.dense_mobility(r, a, 1.0, wall))
    assert np.abs(M - M.T).max() <= 1e-15 * np.abs(M).max()
    assert np.linalg.eigvalsh(M).min() > 0


def test_direct_lower_triangle_matches_mirrored_transpose(orc):
    """The reference mirrors block (i,j)^T into (j,i) (:449-452).  A GPU kernel evaluates
    every ordered pair directly with the SOURCE blob's height instead; both must agree."""
    s, ref, r, a = _suspension(orc, 3, 12, True)
    rf = r.reshape(-1)
    worst = 0.0
    for i in range(0, r.shape[0], 5):
        for j in range(i + 1, r.shape[0], 3):
            up = orc.pair_block(rf, i, j, a, True)
            swapped = rf.reshape(-1, 3)[[j, i]].reshape(-1)
            direct = orc.pair_block(swapped, 0, 1, a, True)  # block (j,i) evaluated with h_i
            worst = max(worst, np.abs(direct - up.T).max() / np.abs(up).max())
    assert worst < 1e-14


def test_dense_and_matrix_free_agree(orc):
    s, ref, r, a = _suspension(orc, 4, 12, True)
    F = np.random.default_rng(0).standard_normal(r.size)
    assert rel_err(orc.apply_M(F, r, a, 0.9, True), orc.apply_M_dense(F, r, a, 0.9, True)) < 1e-14
    rows = np.array([0, 7, 20, 47])
    full = orc.apply_M(F, r, a, 0.9, True).reshape(-1, 3)
    assert np.array_equal(orc.apply_M(F, r, a, 0.9, True, rows=rows).reshape(-1, 3), full[rows])


def test_single_sphere_wall_mobility_analytic(orc):
    """mu_par/mu_0 = 1 - 9/(16h) + 1/(8h^3) - 1/(16h^5),
    mu_perp/mu_0 = 1 - 9/(8h) + 1/(2h^3) - 1/(8h^5) (Swan & Brady), h = z/a."""
    a, eta = 0.37, 1.3
    mu0 = 1.0 / (6 * np.pi * eta * a)
    for h in (1.1, 2.0, 5.0, 40.0):
        M = np.asarray(orc.dense_mobility(np.array([0.3, -0.2, h * a]), a, eta, True))
        par = mu0 * (1 - 9 / (16 * h) + 1 / (8 * h**3) - 1 / (16 * h**5))
        perp = mu0 * (1 - 9 / (8 * h) + 1 / (2 * h**3) - 1 / (8 * h**5))
        assert np.allclose([M[0, 0], M[1, 1], M[2, 2]], [par, par, perp], rtol=1e-13)
        assert np.abs(M - np.diag(np.diag(M))).max() == 0


def test_far_field_is_oseen(orc):
    a, eta = 0.1, 1.0
    d = np.array([300.0, -200.0, 150.0])
    M = np.asarray(orc.dense_mobility(np.concatenate([d, np.zeros(3)]), a, eta, False))[:3, 3:]
    rr = np.linalg.norm(d)
    oseen = (np.eye(3) + np.outer(d, d) / rr**2) / (8 * np.pi * eta * rr)
    assert np.abs(M - oseen).max() / np.abs(oseen).max() < 1e-6


def test_rpy_continuous_at_contact(orc):
    a = 0.5
    e = np.array([0.6, 0.0, 0.8])
    lo = np.asarray(orc.dense_mobility(np.concatenate([e * 2 * a * (1 - 1e-9), np.zeros(3)]), a, 1.0, False))[:3, 3:]
    hi = np.asarray(orc.dense_mobility(np.concatenate([e * 2 * a * (1 + 1e-9), np.zeros(3)]), a, 1.0, False))[:3, 3:]
    assert np.abs(lo - hi).max() < 1e-8 * np.abs(hi).max()


def test_damping_matrix(orc):
    r = np.array([[0, 0, 2.0], [0, 0, 0.25], [0, 0, 0.5]])
    assert np.array_equal(orc.damp_diag(r, 0.5), np.repeat([1.0, 0.5, 1.0], 3))


def test_blob_positions_match_scipy_rotation(orc):
    rng = np.random.default_rng(3)
    from rigid_body_light_b200.shells import icosphere_shell

    _, cfg = icosphere_shell(12)
    X = rng.uniform(-10, 10, (5, 3))
    Q = orc.normalize_quats(rng.standard_normal((5, 4)))
    pos = orc.blob_positions(X, Q, orc.remove_mean(cfg))
    want = np.concatenate([Rotation.from_quat(Q[b], scalar_first=True).apply(cfg) + X[b] for b in range(5)])
    assert np.allclose(pos, want, atol=1e-12)


def test_K_and_KT_are_adjoint_and_match_dense(orc):
    s, ref, r, a = _suspension(orc, 3, 12, False)
    rng = np.random.default_rng(4)
    U, lam = rng.standard_normal(18), rng.standard_normal(r.size)
    K = orc.K_dense(r, s["X"], 12)
    assert np.allclose(orc.K_dot(U, r, s["X"], 12), K @ U, atol=1e-13)
    assert np.allclose(orc.KT_dot(lam, r, s["X"], 12), K.T @ lam, atol=1e-13)
    assert abs(lam @ orc.K_dot(U, r, s["X"], 12) - U @ orc.KT_dot(lam, r, s["X"], 12)) < 1e-12


def test_Kinv_is_left_inverse_of_K(orc):
    s, ref, r, a = _suspension(orc, 3, 12, False)
    K = orc.K_dense(r, s["X"], 12)
    Kinv = orc.Kinv_dense(r, s["X"], s["Q"], ref)
    assert np.abs(Kinv @ K - np.eye(18)).max() < 1e-12


@pytest.mark.parametrize("block", [False, True])
@pytest.mark.parametrize("wall", [False, True])
def test_pc_inverts_its_own_saddle_matrix(orc, wall, block):
    """test_PC (c_rigid_obj.cpp:569-587): apply_PC([Mt lam - K U ; -K^T lam]) = [lam ; U]."""
    s, ref, r, a = _suspension(orc, 3, 12, wall)
    pc = orc.PC(s["X"], s["Q"], ref, a, 1.1, wall, block)
    rng = np.random.default_rng(5)
    lam, U = rng.standard_normal(r.size), rng.standard_normal(18)
    K = orc.K_dense(r, s["X"], 12)
    Mt = np.zeros((r.size, r.size))
    for b in range(3):
        Mt[36 * b:36 * b + 36, 36 * b:36 * b + 36] = np.linalg.inv(pc.invM[b])
    out = pc.apply(np.concatenate([Mt @ lam - K @ U, -K.T @ lam]))
    assert rel_err(out, np.concatenate([lam, U])) < 1e-10


def test_block_pc_is_exact_for_one_body(orc):
    """With one body the block PC's Mt is the full M: PC o [M -K; -K^T 0] = identity."""
    s, ref, r, a = _suspension(orc, 1, 12, True)
    pc = orc.PC(s["X"], s["Q"], ref, a, 1.0, True, True)
    rng = np.random.default_rng(6)
    x = rng.standard_normal(r.size + 6)
    sad = orc.apply_saddle(x, s["X"], s["Q"], ref, a, 1.0, True)
    sad[r.size:] *= -1  # apply_saddle returns +K^T lam, the PC inverts the -K^T row
    assert rel_err(pc.apply(sad), x) < 1e-9


def test_quaternion_update(orc):
    X = np.zeros((1, 3))
    Q = np.array([[1.0, 0, 0, 0]])
    Xn, Qn = orc.evolve(X, Q, np.array([1.0, 2.0, 3.0, 0, 0, np.pi / 2]), 1.0)
    assert np.allclose(Xn, [[1, 2, 3]])
    assert np.allclose(Qn, [[np.cos(np.pi / 4), 0, 0, np.sin(np.pi / 4)]])
    Xn, Qn = orc.evolve(X, Q, np.zeros(6), 1.0)  # theta <= 1e-10 branch (:684)
    assert np.array_equal(Qn, Q)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_golden_cases_reproduce(orc, name):
    """The committed fixtures are what the oracle computes today (guards both)."""
    g = load_golden(name)
    a, eta, wall = float(g["a"]), float(g["eta"]), bool(g["wall"])
    ref = orc.remove_mean(g["cfg"])
    Qn = orc.normalize_quats(g["Q"])
    r = orc.blob_positions(g["X"], Qn, ref)
    n_blb = ref.shape[0]
    assert np.allclose(r, g["r"], atol=1e-14)
    assert rel_err(orc.apply_M(g["lam"], r, a, eta, wall), g["MF"]) < 1e-14
    assert rel_err(g["MF_dense"], g["MF"]) < 1e-13
    assert rel_err(orc.K_dot(g["U"], r, g["X"], n_blb), g["KU"]) < 1e-14
    assert rel_err(orc.KT_dot(g["lam"], r, g["X"], n_blb), g["KTlam"]) < 1e-14
    assert rel_err(orc.apply_saddle(g["vec"], g["X"], Qn, ref, a, eta, wall), g["saddle"]) < 1e-14
    if "pc_diag" in g:
        assert rel_err(orc.PC(g["X"], Qn, ref, a, eta, wall, False).apply(g["vec"]), g["pc_diag"]) < 1e-12
        assert rel_err(orc.PC(g["X"], Qn, ref, a, eta, wall, True).apply(g["vec"]), g["pc_block"]) < 1e-10


@pytest.mark.parametrize("name", ["case_touch_wall", "case_touch_free"])
def test_bd_golden_reproduces(orc, name):
    """The committed BD-step fixture (tests/golden/bd_golden.npz) is what the dense oracle computes
    today, for both Brownian-increment routes; the two routes give different vectors of the same
    covariance, hence different (equally valid) steps."""
    g, bd = load_golden(name), load_golden("bd_golden")
    a, eta, wall, dt = float(g["a"]), float(g["eta"]), bool(g["wall"]), float(g["dt"])
    ref = orc.remove_mean(g["cfg"])
    W, F, kBT = bd[f"{name}/W"], bd[f"{name}/F"], float(bd[f"{name}/kBT"])
    for mode in ("symmetric", "block_cholesky"):
        U, Xn, Qn = orc.bd_step(g["X"], g["Qn"], ref, a, eta, dt, kBT, wall, F, None, *W, noise=mode)
        assert rel_err(U, bd[f"{name}/{mode}/U"]) < 1e-9
        assert rel_err(Xn, bd[f"{name}/{mode}/X"]) < 1e-12 and rel_err(Qn, bd[f"{name}/{mode}/Q"]) < 1e-12
    assert rel_err(bd[f"{name}/symmetric/U"], bd[f"{name}/block_cholesky/U"]) > 1e-3


def test_philox_known_answers_and_normal_statistics(orc):
    """oracle.philox4x32_10 against the Random123 known-answer vectors, and the moments of the
    Box-Muller triplets built on it (the device generator of rbl_bd_step_seeded is checked against
    this in tests/test_gpu_rigid.py)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in kat:
        got = orc.philox4x32_10([np.array([c]) for c in ctr], key)
        assert tuple(int(g[0]) for g in got) == want
    n = 400000
    W = orc.philox_normals(seed=42, step=3, first=10**10, n=n)  # a 64-bit element offset too
    for w in W:
        assert abs(w.mean()) < 4 / np.sqrt(n) and abs(w.var() - 1) < 4 * np.sqrt(2 / n)
        assert abs((w ** 4).mean() - 3) < 0.1
    c = np.corrcoef(np.stack(W))
    assert np.abs(c - np.eye(3)).max() < 4 / np.sqrt(n)
    # pure function of (seed, step, element): a shifted window reproduces the overlap, other keys do not
    V = orc.philox_normals(seed=42, step=3, first=10**10 + 1000, n=500)
    assert all(np.array_equal(v, w[1000:1500]) for v, w in zip(V, W))
    assert not np.array_equal(orc.philox_normals(42, 4, 10**10, 10)[0], W[0][:10])
    assert not np.array_equal(orc.philox_normals(43, 3, 10**10, 10)[0], W[0][:10])
