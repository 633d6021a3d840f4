"""The thermal drift of the BD step, checked deterministically on the CPU oracle.

The overdamped Langevin equation of rigid bodies needs the drift kBT div_Q(N), N = (K^T M^-1 K)^-1.
The step (include/rbl.h rbl_bd_step) produces it from two pieces: the
midpoint evaluation of A = N K^T M^-1 along the predictor displacement K^-1 M^{1/2} W_1, and the
random finite difference of M in the slip row.  A third candidate, the random finite difference
of K^T in the force row (the reference has KT_RFD_from_U, c_rigid_obj.cpp:842-863), has zero
expectation with this predictor.  All expectations are quadratic forms in the noise, so they are
summed exactly over basis vectors instead of sampled; the sum must equal div_Q(N) computed by
finite differences of N itself.  Done for a bent, chiral body next to the wall and for an
icosahedral shell."""
import numpy as np
import pytest


def _setup(orc, cfg, X, Q, a, eta):
    ref = orc.remove_mean(cfg)
    n_blb = ref.shape[0]

    def at(X_, Q_):
        r = orc.blob_positions(X_, Q_, ref)
        M = np.asarray(orc.dense_mobility(r, a, eta, True))
        B = orc.damp_diag(r, a)
        assert np.all(B == 1.0)  # stay clear of the wall-overlap layer: B M B = M
        K = orc.K_dense(r, X_, n_blb)
        Kinv = orc.Kinv_dense(r, X_, Q_, ref)
        Minv = np.linalg.inv(M)
        N = np.linalg.inv(K.T @ Minv @ K)
        return {"M": M, "K": K, "Kinv": Kinv, "N": N, "A": N @ K.T @ Minv}

    def moved(v6):
        return orc.update_X_Q(X, Q, v6)

    return at, moved


def _ddir(at, moved, key, v6, h=2e-6):
    """directional derivative of at(.)[key] along the rigid displacement v6 (central difference)"""
    p, m = at(*moved(h * v6)), at(*moved(-h * v6))
    return (p[key] - m[key]) / (2 * h)


def _terms(orc, cfg, X, Q, a, eta):
    at, moved = _setup(orc, cfg, X, Q, a, eta)
    base = at(X, Q)
    M, K, Kinv, N, A = (base[k] for k in ("M", "K", "Kinv", "N", "A"))
    nd = K.shape[1]  # 6 rigid degrees of freedom per body
    assert np.allclose(Kinv @ K, np.eye(nd), atol=1e-12)
    n3 = M.shape[0]
    E = np.eye(nd)
    div_N = sum(_ddir(at, moved, "N", E[k])[:, k] for k in range(nd))
    # (1) noise through A(q') with q' - q = sqrt(kBT dt) K^-1 g, g ~ N(0, M):  E = d_k(A) M K^-T e_k
    t_mid = sum(_ddir(at, moved, "A", E[k]) @ (M @ Kinv[k]) for k in range(nd))
    # (2) slip-row RFD of M: E[(M(q+) - M(q-)) W / delta], q+- = q +- (delta/2) K^-1 W, summed over W = e_j
    rfd_M = sum(_ddir(at, moved, "M", Kinv[:, j])[:, j] for j in range(n3))
    # (3) force-row RFD of K^T, same displacements
    rfd_KT = sum(_ddir(at, moved, "K", Kinv[:, j]).T[:, j] for j in range(n3))
    return div_N, t_mid, A @ rfd_M, N @ rfd_KT, N


def _random_quat(seed):
    q = np.random.default_rng(seed).standard_normal(4)
    return (q / np.linalg.norm(q))[None, :]


def test_drift_of_an_anisotropic_body_next_to_the_wall(orc):
    cfg = np.array([[0.0, 0, 0], [0.7, 0, 0], [1.4, 0, 0], [0, 0.8, 0], [0.3, 0.2, 0.9]])  # a bent, chiral 5-blob body
    X, Q = np.array([[0.3, -0.2, 2.1]]), _random_quat(4)
    div_N, t_mid, t_M, t_KT, N = _terms(orc, cfg, X, Q, a=0.3, eta=1.0)
    scale = np.abs(N).max()
    assert np.abs(div_N).max() > 1e-3 * scale                      # there IS a drift to get right
    assert np.abs(t_mid + t_M - div_N).max() < 1e-6 * scale          # midpoint + RFD give exactly kBT div N
    assert np.abs(t_KT).max() < 1e-8 * scale                       # the K^T finite difference has zero mean
    assert np.abs(t_mid - div_N).max() > 1e-3 * scale                # without the RFD of M the drift is wrong
    assert np.abs(t_M - div_N).max() > 1e-3 * scale                  # and so it is without the midpoint


def test_drift_of_two_hydrodynamically_coupled_bodies(orc):
    """two copies of the bent body, one above the other's shoulder: the 12 x 12 body mobility N
    couples them, and the identity must hold for the whole vector, cross terms included"""
    cfg = np.array([[0.0, 0, 0], [0.7, 0, 0], [1.4, 0, 0], [0, 0.8, 0], [0.3, 0.2, 0.9]])
    X = np.array([[0.3, -0.2, 2.1], [1.9, 0.8, 3.0]])
    Q = np.concatenate([_random_quat(4), _random_quat(9)])
    div_N, t_mid, t_M, t_KT, N = _terms(orc, cfg, X, Q, a=0.3, eta=1.0)
    scale = np.abs(N).max()
    assert np.abs(N[:6, 6:]).max() > 1e-2 * scale                  # the bodies do interact
    assert np.abs(t_mid + t_M - div_N).max() < 1e-6 * scale
    assert np.abs(t_KT).max() < 1e-8 * scale
    assert np.abs(t_mid - div_N).max() > 1e-3 * scale


def test_kt_drift_vanishes_for_an_icosahedral_shell(orc):
    from rigid_body_light_b200.shells import icosphere_shell

    params, cfg = icosphere_shell(12)
    X, Q = np.array([[0.0, 0.0, 1.6]]), _random_quat(7)
    div_N, t_mid, t_M, t_KT, N = _terms(orc, cfg, X, Q, a=params["sep"] / 2, eta=1.0)
    scale = np.abs(N).max()
    assert np.abs(t_KT).max() < 1e-8 * scale
    assert np.abs(t_mid + t_M - div_N).max() < 1e-6 * scale
    assert abs(div_N[2]) > 1e-3 * scale  # the familiar d(mu_perp)/dh drift away from the wall


def test_oracle_rfd_estimators_have_these_expectations(orc):
    """the RFD lines of oracle.bd_step, averaged EXACTLY over W_r = e_j, give (d_k M) K^-T e_k and
    (d_k K^T) K^-T e_k"""
    cfg = np.array([[0.0, 0, 0], [0.7, 0, 0], [1.4, 0, 0], [0, 0.8, 0], [0.3, 0.2, 0.9]])
    X, Q = np.array([[0.3, -0.2, 2.1]]), _random_quat(4)
    a, eta, delta = 0.3, 1.0, 1e-5
    ref = orc.remove_mean(cfg)
    r = orc.blob_positions(X, Q, ref)
    n3 = r.size
    rfd_M, rfd_KT = np.zeros(n3), np.zeros(6)
    for j in range(n3):
        Wr = np.zeros(n3)
        Wr[j] = 1.0
        uom = orc.Kinv_apply(Wr, r, X, Q, ref)
        Xp, Qp = orc.update_X_Q(X, Q, 0.5 * delta * uom)
        Xn, Qn = orc.update_X_Q(X, Q, -0.5 * delta * uom)
        rp, rn = orc.blob_positions(Xp, Qp, ref), orc.blob_positions(Xn, Qn, ref)
        rfd_M += (orc.apply_M(Wr, rp, a, eta, True) - orc.apply_M(Wr, rn, a, eta, True)) / delta
        rfd_KT += (orc.KT_dot(Wr, rp, Xp, ref.shape[0]) - orc.KT_dot(Wr, rn, Xn, ref.shape[0])) / delta
    at, moved = _setup(orc, cfg, X, Q, a, eta)
    Kinv = at(X, Q)["Kinv"]
    E = np.eye(6)
    want_M = sum(_ddir(at, moved, "M", E[k]) @ Kinv[k] for k in range(6))
    want_KT = sum(_ddir(at, moved, "K", E[k]).T @ Kinv[k] for k in range(6))
    assert np.abs(rfd_M - want_M).max() < 1e-5 * np.abs(want_M).max()
    assert np.abs(rfd_KT - want_KT).max() < 1e-6 and np.abs(want_KT).max() < 1e-8  # both vanish


def test_oracle_preconditioned_noise_has_the_covariance_of_the_mobility(orc):
    """S = L (G A G^T)^{1/2} satisfies S S^T = A for both factor constructions (per-body Cholesky
    with the wall, one rotated reference factor in free space)."""
    from rigid_body_light_b200.shells import sphere_suspension

    for wall in (True, False):
        s = sphere_suspension(3, 12, wall)
        ref = orc.remove_mean(s["cfg"])
        r = orc.blob_positions(s["X"], s["Q"], ref)
        A = np.asarray(orc.dense_mobility(r, s["a"], 1.0, wall))
        fac = orc.noise_factors(r, s["Q"], ref, s["a"], 1.0, wall)
        for b, Lb in enumerate(fac):
            sl = slice(36 * b, 36 * b + 36)
            assert np.allclose(Lb @ Lb.T, A[sl, sl], rtol=0, atol=1e-13)
        S = np.stack([orc.noise_block_cholesky(fac, A, e) for e in np.eye(A.shape[0])], axis=1)
        assert np.linalg.norm(S @ S.T - A) / np.linalg.norm(A) < 1e-10
