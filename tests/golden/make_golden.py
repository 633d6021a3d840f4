"""Generates the committed golden fixtures under tests/golden/.

Run HERE (the build container), where /root/reference exists:

    python tests/golden/make_golden.py

* ``pair_golden.npz``  -- outputs of the REFERENCE's own pair kernels (mobilityUFRPY,
  mobilityUFSingleWallCorrection, /root/reference/src/c_rigid_obj.cpp:31-142) compiled
  from the reference source by oracle/build_ref.sh, on seeded random arguments, in
  float64 and float32.  These pin the oracle's restatement wherever the tests run
  (the GPU box has no /root/reference).
* ``shells_check.npz`` -- nearest-neighbour distance of every reference CSV blob to the
  regenerated icosphere shells (proves rigid_body_light_b200.shells reproduces
  structures/shell_N_*.csv as point sets).
* ``case_*.npz``       -- seeded small suspensions with the float64 oracle's outputs for
  every operator of the hot path (apply_M, K, K^T, Kinv, both PCs, saddle, evolve).
  The reference itself holds no golden values for these.
* ``apply_M_ref_golden.npz`` -- apply_M of the cases above (and of a ragged cloud) computed by the
  REFERENCE'S OWN assembly + apply_M members compiled from the reference source
  (oracle/_ref/libref_apply_M.so); ``--apply-M-ref-only`` regenerates just this file.
* ``members_ref_golden.npz`` -- quaternion normalisation, blob placement, K, K^T, K^-1, K^-T, both
  preconditioners and the integrator of the cases above through the REFERENCE'S OWN member functions
  (oracle/_ref/libref_members.so).
* ``bd_golden.npz``    -- one fluctuating BD step per touching-sphere case with fixed noise through
  the dense oracle (both Brownian-increment routes).  ``--bd-only`` regenerates just this file.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import oracle as orc  # noqa: E402
from rigid_body_light_b200 import shells  # noqa: E402


def pair_golden():
    R = orc.ref_pair_lib()
    assert R is not None, "oracle/_ref/libref_pair.so missing: run make -C oracle"
    rng = np.random.default_rng(12345)
    n = 400
    # separations spanning both RPY branches (r/a in (0.05, 12)) and wall heights
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= rng.uniform(0.05, 12.0, (n, 1))
    hi = rng.uniform(0.05, 6.0, n)
    hj = rng.uniform(0.05, 6.0, n)
    out = {"d": d, "hi": hi, "hj": hj}
    for sfx, ct, dt in (("f64", ctypes.c_double, np.float64), ("f32", ctypes.c_float, np.float32)):
        rpy = np.zeros((n, 6), dt)
        wall = np.zeros((n, 9), dt)
        wself = np.zeros((n, 9), dt)
        for k in range(n):
            dk = d[k].astype(dt)
            m6 = np.zeros(6, dt)
            getattr(R, f"ref_rpy_pair_{sfx}")(ct(dk[0]), ct(dk[1]), ct(dk[2]), m6.ctypes.data, 0, 1, ct(1.0))
            rpy[k] = m6
            m9 = np.zeros(9, dt)
            # arguments as rotne_prager_tensor passes them (:441-444) with a = 1:
            # (rx, ry, rz + 2 z_j, h_j) and rz = z_i - z_j
            Rz = dt(hi[k]) + dt(hj[k])
            st = getattr(R, f"ref_wall_pair_{sfx}")(ct(dk[0]), ct(dk[1]), ct(Rz), m9.ctypes.data, 0, 1, ct(dt(hj[k])))
            assert st == 0
            wall[k] = m9
            m9 = np.zeros(9, dt)
            st = getattr(R, f"ref_wall_pair_{sfx}")(ct(0), ct(0), ct(2 * dt(hj[k])), m9.ctypes.data, 3, 3, ct(dt(hj[k])))
            assert st == 0
            wself[k] = m9
        out[f"rpy_{sfx}"] = rpy
        out[f"wall_{sfx}"] = wall
        out[f"wall_self_{sfx}"] = wself
    # the throw: hj < 0
    m9 = np.zeros(9)
    out["below_status"] = np.array(R.ref_wall_pair_f64(ctypes.c_double(0.1), ctypes.c_double(0.2), ctypes.c_double(1.0), m9.ctypes.data, 0, 1, ctypes.c_double(-0.5)))
    np.savez_compressed(os.path.join(HERE, "pair_golden.npz"), **out)
    print("pair_golden.npz:", n, "pairs x {f64,f32}")


def shells_check():
    from scipy.spatial import cKDTree
    out = {}
    for n in shells.SHELLS:
        fn = f"/root/reference/structures/shell_N_{n}.csv"
        p_ref, c_ref = shells.load_config(fn)
        p, c = shells.icosphere_shell(n)
        d, idx = cKDTree(c).query(c_ref)
        assert len(set(idx.tolist())) == n
        out[f"maxdist_{n}"] = np.array(d.max())
        out[f"sep_{n}"] = np.array([p_ref["sep"], p["sep"]])
        print(f"shell_N_{n}: max |generated - csv| = {d.max():.2e} (csv prints 8 decimals)")
    np.savez_compressed(os.path.join(HERE, "shells_check.npz"), **out)


CASES = {
    # name: (n_bodies, shell, wall, a or None (= sep/2), eta, z_shift)
    "case_overlap_wall": (3, 12, True, 1.0, 1.0, 1.0),   # reference test regime: a=1 >> sep/2
    "case_overlap_free": (3, 12, False, 1.0, 1.0, 0.0),
    "case_touch_wall": (4, 42, True, None, 0.7, 0.0),    # benchmark regime: a = sep/2
    "case_touch_free": (4, 42, False, None, 0.7, 0.0),
    "case_near_wall": (5, 12, True, 0.3, 1.3, -0.6),    # blobs inside the B-damping layer z < a
}


def make_case(name):
    nb, shell, wall, a, eta, zshift = CASES[name]
    s = shells.sphere_suspension(nb, shell, wall)
    a = a if a is not None else s["a"]
    X = s["X"].copy()
    X[:, 2] += zshift
    Q = s["Q"] * np.linspace(0.5, 2.0, nb)[:, None]  # un-normalised on purpose (:216)
    ref = orc.remove_mean(s["cfg"])
    Qn = orc.normalize_quats(Q)
    r = orc.blob_positions(X, Qn, ref)
    n_blb = ref.shape[0]
    rng = np.random.default_rng(2)
    lam = rng.standard_normal(r.size)
    U = rng.standard_normal(6 * nb)
    vec = rng.standard_normal(r.size + 6 * nb)
    dt = 0.01
    out = dict(cfg=s["cfg"], X=X, Q=Q, a=np.array(a), eta=np.array(eta), wall=np.array(wall), dt=np.array(dt),
               lam=lam, U=U, vec=vec, Qn=Qn, r=r)
    out["MF"] = orc.apply_M(lam, r, a, eta, wall)
    out["MF_dense"] = orc.apply_M_dense(lam, r, a, eta, wall)
    out["KU"] = orc.K_dot(U, r, X, n_blb)
    out["KTlam"] = orc.KT_dot(lam, r, X, n_blb)
    Kinv = orc.Kinv_dense(r, X, Qn, ref)
    out["Kinv_lam"] = Kinv @ lam
    out["KinvT_U"] = Kinv.T @ U
    out["saddle"] = orc.apply_saddle(vec, X, Qn, ref, a, eta, wall)
    if r[:, 2].min() >= a or not wall:
        # (inside the overlap layer z < a the wall-corrected self mobility goes negative and
        #  the reference's LLT of K^T Mt^-1 K breaks down too: no PC golden for that case)
        out["pc_diag"] = orc.PC(X, Qn, ref, a, eta, wall, False).apply(vec)
        out["pc_block"] = orc.PC(X, Qn, ref, a, eta, wall, True).apply(vec)
    Xe, Qe = orc.evolve(X, Qn, U, dt)
    out["X_evolved"], out["Q_evolved"] = Xe, Qe
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "N =", r.shape[0], "zmin/a =", r[:, 2].min() / a)


def apply_M_ref_golden():
    """apply_M of every committed case computed by THE REFERENCE'S OWN CODE (rotne_prager_tensor +
    make_damp_mat + apply_M, c_rigid_obj.cpp:413-459,618-659, compiled from the reference source by
    oracle/build_ref.sh with oracle/eigen_shim.inc standing in for Eigen3), in double and in float,
    plus a ragged random cloud with overlapping and touching pairs.  This is what pins the oracle's
    M.F -- and through it the CUDA product -- to the reference where /root/reference is absent."""
    assert orc.ref_apply_M_lib() is not None, "oracle/_ref/libref_apply_M.so missing: run make -C oracle"
    out = {}
    for name in CASES:
        g = dict(np.load(os.path.join(HERE, name + ".npz")))
        a, eta, wall = float(g["a"]), float(g["eta"]), bool(g["wall"])
        out[f"{name}/f64"] = orc.ref_apply_M(g["lam"], g["r"], a, eta, wall)
        out[f"{name}/f32"] = orc.ref_apply_M(g["lam"].astype(np.float32), g["r"].astype(np.float32), a, eta, wall, dtype=np.float32)
    rng = np.random.default_rng(77)
    n, a, eta = 257, 0.11, 0.9
    r = rng.uniform(0.05, 3.0, (n, 3))
    r[1] = r[0] + [0.7 * a, 0, 0.1 * a]
    r[3] = r[2] + [0, 2.0 * a, 0]
    F = rng.standard_normal(3 * n)
    out["cloud/r"], out["cloud/F"], out["cloud/a"], out["cloud/eta"] = r, F, np.array(a), np.array(eta)
    for wall in (False, True):
        out[f"cloud/wall{int(wall)}/f64"] = orc.ref_apply_M(F, r, a, eta, wall)
        out[f"cloud/wall{int(wall)}/f32"] = orc.ref_apply_M(F.astype(np.float32), r.astype(np.float32), a, eta, wall, dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "apply_M_ref_golden.npz"), **out)
    print("apply_M_ref_golden.npz:", len(out), "arrays from the reference's own apply_M")


def members_ref_golden():
    """Every committed case through THE REFERENCE'S OWN member functions (oracle.RefBody =
    oracle/_ref/libref_members.so: setParameters / setConfig / set_K_mats / multi_body_pos / K_x_U /
    KT_x_Lam / Kinv_x_V / KTinv_x_F / apply_PC (diagonal and block) / evolve_X_Q compiled from the
    reference source), in double and float.  Pins the oracle's numpy restatement of those members where /root/reference is absent."""
    out = {}
    for name in CASES:
        g = dict(np.load(os.path.join(HERE, name + ".npz")))
        a, eta, wall, dt = float(g["a"]), float(g["eta"]), bool(g["wall"]), float(g["dt"])
        for sfx, ndt in (("f64", np.float64), ("f32", np.float32)):
            rb = orc.RefBody(g["cfg"], g["X"], g["Q"], a, eta, dt, wall_PC=wall, dtype=ndt)
            X, Q = rb.get_config()
            out[f"{name}/{sfx}/Qn"], out[f"{name}/{sfx}/r"] = Q, rb.positions()
            out[f"{name}/{sfx}/KU"], out[f"{name}/{sfx}/KTlam"] = rb.K_dot(g["U"]), rb.KT_dot(g["lam"])
            out[f"{name}/{sfx}/Kinv_lam"], out[f"{name}/{sfx}/KinvT_U"] = rb.Kinv_dot(g["lam"]), rb.KTinv_dot(g["U"])
            for blk in (False, True):  # both preconditioners (lazily built at the first apply_PC, :591-596)
                pcb = orc.RefBody(g["cfg"], g["X"], g["Q"], a, eta, dt, wall_PC=wall, block_PC=blk, dtype=ndt)
                out[f"{name}/{sfx}/pc_{'block' if blk else 'diag'}"] = pcb.apply_PC(g["vec"])
            rb.evolve(g["U"])
            Xe, Qe = rb.get_config()
            out[f"{name}/{sfx}/X_evolved"], out[f"{name}/{sfx}/Q_evolved"] = Xe, Qe
            out[f"{name}/{sfx}/KU_evolved"] = rb.K_dot(g["U"])  # K rebuilt by evolve_X_Q (:876)
    np.savez_compressed(os.path.join(HERE, "members_ref_golden.npz"), **out)
    print("members_ref_golden.npz:", len(out), "arrays from the reference's own member functions")


def bd_golden():
    """One fluctuating BD step per touching-sphere case with FIXED noise, through the dense oracle:
    Brownian increments by both routes (symmetric square root; block-Cholesky preconditioned root)
    and the resulting rigid velocities / new configuration.  Pins oracle.bd_step and gives the
    GPU tests a committed target for rbl_bd_step / rbl_lanczos_sqrt."""
    out = {}
    for name in ("case_touch_wall", "case_touch_free"):
        g = dict(np.load(os.path.join(HERE, name + ".npz")))
        a, eta, wall, dt = float(g["a"]), float(g["eta"]), bool(g["wall"]), float(g["dt"])
        ref = orc.remove_mean(g["cfg"])
        nb, n3 = g["X"].shape[0], g["r"].size
        rng = np.random.default_rng(31)
        F = rng.standard_normal(6 * nb)
        W = [rng.standard_normal(n3) for _ in range(3)]
        kBT = 0.004
        M_raw = np.asarray(orc.dense_mobility(g["r"], a, eta, wall))
        A = M_raw
        if wall:
            B = orc.damp_diag(g["r"], a)
            A = B[:, None] * M_raw * B[None, :]
        from scipy.linalg import sqrtm
        out[f"{name}/F"], out[f"{name}/W"], out[f"{name}/kBT"] = F, np.array(W), np.array(kBT)
        out[f"{name}/noise_symmetric"] = np.real(sqrtm(A)) @ W[0]
        out[f"{name}/noise_block_cholesky"] = orc.noise_block_cholesky(orc.noise_factors(g["r"], g["Qn"], ref, a, eta, wall), A, W[0])
        for mode in ("symmetric", "block_cholesky"):
            U, Xn, Qn = orc.bd_step(g["X"], g["Qn"], ref, a, eta, dt, kBT, wall, F, None, *W, noise=mode)
            out[f"{name}/{mode}/U"], out[f"{name}/{mode}/X"], out[f"{name}/{mode}/Q"] = U, Xn, Qn
        print("bd_golden", name, "|U| =", np.linalg.norm(U))
    np.savez_compressed(os.path.join(HERE, "bd_golden.npz"), **out)


if __name__ == "__main__":
    if "--bd-only" in sys.argv:
        bd_golden()
        sys.exit(0)
    if "--apply-M-ref-only" in sys.argv:
        apply_M_ref_golden()
        members_ref_golden()
        sys.exit(0)
    pair_golden()
    shells_check()
    for c in CASES:
        make_case(c)
    apply_M_ref_golden()
    members_ref_golden()
    bd_golden()
