"""CPU oracle for the Rigid_Body_Light hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``rigid_body_light_b200/`` or ``Rigid/`` may import this module.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, as the checker or the reported CPU
baseline, never as the product path.

What it restates (citations are into /root/reference/src/):

* the O(N^2) pair arithmetic lives in ``rbl_oracle.c`` (c_rigid_obj.cpp:31-142,
  413-459, 618-659) and is reached through ctypes;
* everything O(N) around it is plain numpy float64 below, one function per
  reference member function, each citing the lines it follows.

Pinning status: every function here is pinned against the reference's OWN code, compiled
from the reference source where it lies by build_ref.sh into ``oracle/_ref`` (Eigen3 is
absent: ``eigen_shim.inc`` supplies the subset those members use) -- the two pair kernels
and the dense ``apply_M`` bit for bit in float and double, state handling, placement, K,
K^T, K^-1, both preconditioners and the integrator to rounding (tests/test_oracle_vs_ref.py,
live through ``RefBody`` where ``_ref`` exists, and against committed outputs of the
reference's code under tests/golden/ everywhere).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("rbl_oracle.c", "oracle_impl.inc")]
    stale = force or not os.path.exists(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in src
    )
    need_ref = os.path.exists("/root/reference/src/c_rigid_obj.cpp") and not all(
        os.path.exists(os.path.join(_HERE, "_ref", f)) for f in ("libref_pair.so", "libref_members.so")
    )
    if stale or need_ref:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
        _declare(_LIB)
    return _LIB


def ref_pair_lib():
    """The reference's own pair kernels (oracle/_ref/libref_pair.so) or None."""
    global _REF
    if _REF is None:
        p = os.path.join(_HERE, "_ref", "libref_pair.so")
        if not os.path.exists(p):
            build()
        if not os.path.exists(p):
            return None
        _REF = ctypes.CDLL(p)
        for sfx, ct in (("f64", ctypes.c_double), ("f32", ctypes.c_float)):
            f = getattr(_REF, f"ref_rpy_pair_{sfx}")
            f.argtypes = [ct, ct, ct, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ct]
            f.restype = None
            g = getattr(_REF, f"ref_wall_pair_{sfx}")
            g.argtypes = [ct, ct, ct, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ct]
            g.restype = ctypes.c_int
    return _REF


_REF_APPLY = None


def ref_apply_M_lib():
    """The reference's own member functions on the product path compiled from the reference source
    (oracle/_ref/libref_members.so, see build_ref.sh) or None."""
    global _REF_APPLY
    if _REF_APPLY is None:
        p = os.path.join(_HERE, "_ref", "libref_members.so")
        if not os.path.exists(p):
            build()
        if not os.path.exists(p):
            return None
        L = ctypes.CDLL(p)
        vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        for sfx in ("f64", "f32"):
            getattr(L, f"ref_apply_M_{sfx}").argtypes = [vp, vp, ci, cd, cd, ci, vp]
            getattr(L, f"refm_create_{sfx}").restype = vp
            getattr(L, f"refm_destroy_{sfx}").argtypes = [vp]
            getattr(L, f"refm_destroy_{sfx}").restype = None
            getattr(L, f"refm_set_parameters_{sfx}").argtypes = [vp, cd, cd, cd, cd, vp, ci]
            getattr(L, f"refm_set_flags_{sfx}").argtypes = [vp, ci, ci]
            getattr(L, f"refm_set_config_{sfx}").argtypes = [vp, vp, vp, ci]
            getattr(L, f"refm_get_config_{sfx}").argtypes = [vp, vp, vp]
            getattr(L, f"refm_apply_M_{sfx}").argtypes = [vp, vp, vp, ci, vp]
            for name in ("positions", "evolve"):
                getattr(L, f"refm_{name}_{sfx}").argtypes = [vp, vp]
            getattr(L, f"refm_RHS_{sfx}").argtypes = [vp, vp, vp, vp, vp, vp, cd, vp]
            for name in ("K_x_U", "KT_x_Lam", "Kinv_x_V", "KTinv_x_F", "apply_PC", "M_RFD", "M_half_W", "KTinv_RFD"):
                getattr(L, f"refm_{name}_{sfx}").argtypes = [vp, vp, vp]
            for name in ("M_RFD_from_U", "KT_RFD_from_U", "update_X_Q_out"):
                getattr(L, f"refm_{name}_{sfx}").argtypes = [vp, vp, vp, vp]
            getattr(L, f"refm_M_RFD_cfgs_{sfx}").argtypes = [vp, vp, cd, vp, vp]
            getattr(L, f"refm_evolve_RFD_{sfx}").argtypes = [vp, vp]
            getattr(L, f"refm_set_split_rand_{sfx}").argtypes = [vp, ci]
        _REF_APPLY = L
    return _REF_APPLY


class RefBody:
    """The reference's own CManyBodies members (state handling, placement, K / K^T / K^-1, apply_M,
    both preconditioners, integrator: c_rigid_obj.cpp:176-410,413-567,589-659,678-728,865-878) compiled from the reference
    source with oracle/eigen_shim.inc standing in for the absent Eigen3.  Same call sequence as
    src/Rigid.py: parameters, flags, configuration (setConfig + set_K_mats).  TEST INFRASTRUCTURE."""

    def __init__(self, rigid_config, X, Q, a, eta, dt, wall_PC=False, block_PC=False, dtype=np.float64):
        L = ref_apply_M_lib()
        if L is None:
            raise OracleError("oracle/_ref/libref_members.so is not built (no /root/reference here)")
        self.L, self.dt_ = L, np.dtype(dtype)
        self.sfx = _CT[self.dt_][0]
        self.h = ctypes.c_void_p(getattr(L, f"refm_create_{self.sfx}")())
        cfg = _prep(rigid_config, dtype)
        self.n_blb = cfg.size // 3
        self._call("set_parameters", float(a), float(dt), 1.0, float(eta), cfg.ctypes.data, self.n_blb)
        self._call("set_flags", int(block_PC), int(wall_PC))
        self.set_config(X, Q)

    def _call(self, name, *args):
        st = getattr(self.L, f"refm_{name}_{self.sfx}")(self.h, *args)
        if st != 0:
            raise OracleError(f"reference member {name} threw")

    def __del__(self):
        try:
            getattr(self.L, f"refm_destroy_{self.sfx}")(self.h)
        except Exception:
            pass

    def _vec(self, x):
        return _prep(x, self.dt_)

    def set_config(self, X, Q):
        X, Q = self._vec(X), self._vec(Q)
        self.n_bod = X.size // 3
        self._call("set_config", X.ctypes.data, Q.ctypes.data, self.n_bod)

    def get_config(self):
        X, Q = np.empty(3 * self.n_bod, self.dt_), np.empty(4 * self.n_bod, self.dt_)
        self._call("get_config", X.ctypes.data, Q.ctypes.data)
        return X.reshape(-1, 3), Q.reshape(-1, 4)

    def positions(self):
        out = np.empty(3 * self.n_bod * self.n_blb, self.dt_)
        self._call("positions", out.ctypes.data)
        return out.reshape(-1, 3)

    def _mv(self, name, x, n_out):
        x = self._vec(x)
        out = np.empty(n_out, self.dt_)
        self._call(name, x.ctypes.data, out.ctypes.data)
        return out

    def K_dot(self, U):
        return self._mv("K_x_U", U, 3 * self.n_bod * self.n_blb)

    def KT_dot(self, lam):
        return self._mv("KT_x_Lam", lam, 6 * self.n_bod)

    def Kinv_dot(self, V):
        return self._mv("Kinv_x_V", V, 6 * self.n_bod)

    def KTinv_dot(self, F):
        return self._mv("KTinv_x_F", F, 3 * self.n_bod * self.n_blb)

    def apply_M(self, F, r):
        F, r = self._vec(F), self._vec(r)
        out = np.empty_like(F)
        self._call("apply_M", F.ctypes.data, r.ctypes.data, F.size // 3, out.ctypes.data)
        return out

    def apply_PC(self, b):
        return self._mv("apply_PC", b, 3 * self.n_bod * self.n_blb + 6 * self.n_bod)

    def M_RFD(self, W):
        """the reference's M_RFD (:769-796) with ITS rand_vector replaced by the given noise W"""
        return self._mv("M_RFD", W, 3 * self.n_bod * self.n_blb)

    # the remaining random finite differences and RFD-sized configuration updates, noise injected
    def KTinv_RFD(self, W6):
        """KTinv_RFD (:743-767) with rand_vector(6 N_bod) = W6"""
        return self._mv("KTinv_RFD", W6, 6 * self.n_bod)

    def M_RFD_from_U(self, U, W):
        """M_RFD_from_U (:818-840), delta = 1e-3"""
        U, W = self._vec(U), self._vec(W)
        out = np.empty_like(W)
        self._call("M_RFD_from_U", U.ctypes.data, W.ctypes.data, out.ctypes.data)
        return out

    def KT_RFD_from_U(self, U, W):
        """KT_RFD_from_U (:842-863), delta = 1e-3"""
        U, W = self._vec(U), self._vec(W)
        out = np.empty(6 * self.n_bod, self.dt_)
        self._call("KT_RFD_from_U", U.ctypes.data, W.ctypes.data, out.ctypes.data)
        return out

    def M_RFD_cfgs(self, U, delta):
        """M_RFD_cfgs (:798-816): blob positions of q +- (delta/2) U"""
        U = self._vec(U)
        n3 = 3 * self.n_bod * self.n_blb
        rp, rm = np.empty(n3, self.dt_), np.empty(n3, self.dt_)
        self._call("M_RFD_cfgs", U.ctypes.data, float(delta), rp.ctypes.data, rm.ctypes.data)
        return rp.reshape(-1, 3), rm.reshape(-1, 3)

    def update_X_Q_out(self, U):
        """update_X_Q_out (:712-728): (X, Q [w x y z]) displaced by U, state untouched"""
        U = self._vec(U)
        X, Q = np.empty(3 * self.n_bod, self.dt_), np.empty(4 * self.n_bod, self.dt_)
        self._call("update_X_Q_out", U.ctypes.data, X.ctypes.data, Q.ctypes.data)
        return X.reshape(-1, 3), Q.reshape(-1, 4)

    def set_split_rand(self, on):
        """the class's split_rand member (:150); False selects the single-increment branch (:949-953)"""
        self._call("set_split_rand", int(bool(on)))

    def evolve_RFD(self, U):
        """evolve_X_Q_RFD (:880-893)"""
        U = self._vec(U)
        self._call("evolve_RFD", U.ctypes.data)

    def M_half_W(self, W):
        """the reference's M_half_W (:661-675: chol(B M B) W) with the given noise W"""
        return self._mv("M_half_W", W, 3 * self.n_bod * self.n_blb)

    def RHS_and_Midpoint(self, slip, force, W1, W2, Wr, kBT):
        """the reference's RHS_and_Midpoint (:917-976) with rand_vector returning W1, W2, Wr in the order
        it draws them: [slip - kBT M_RFD - c2 (M^{1/2}W1 - M^{1/2}W2) ; -force]"""
        a = [self._vec(x) for x in (slip, force, W1, W2, Wr)]
        out = np.empty(3 * self.n_bod * self.n_blb + 6 * self.n_bod, self.dt_)
        self._call("RHS", a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, a[4].ctypes.data, float(kBT),
                   out.ctypes.data)
        return out

    def evolve(self, U):
        U = self._vec(U)
        self._call("evolve", U.ctypes.data)


def ref_apply_M(F, r, a, eta, wall, dtype=np.float64):
    """U = apply_M(F, r) computed BY THE REFERENCE'S OWN CODE (c_rigid_obj.cpp:413-459,618-659
    compiled from /root/reference with oracle/eigen_shim.inc standing in for the absent Eigen3).
    Raises OracleError for a blob below the wall, like the reference throws.  None if the library
    is not available."""
    L = ref_apply_M_lib()
    if L is None:
        return None
    F, r = _prep(F, dtype), _prep(r, dtype)
    n = r.size // 3
    sfx, _ = _CT[np.dtype(dtype)]
    U = np.empty(3 * n, dtype=dtype)
    st = getattr(L, f"ref_apply_M_{sfx}")(F.ctypes.data, r.ctypes.data, n, float(a), float(eta), int(wall), U.ctypes.data)
    if st != 0:
        raise OracleError("reference apply_M threw (blob below the wall, c_rigid_obj.cpp:95-97)")
    return U


_CT = {np.dtype(np.float64): ("f64", ctypes.c_double), np.dtype(np.float32): ("f32", ctypes.c_float)}


def _declare(L):
    vp, ci = ctypes.c_void_p, ctypes.c_int
    for sfx, ct in (("f64", ctypes.c_double), ("f32", ctypes.c_float)):
        getattr(L, f"orc_dense_mobility_{sfx}").argtypes = [vp, ci, ct, ct, ci, vp]
        getattr(L, f"orc_apply_M_dense_{sfx}").argtypes = [vp, vp, ci, ct, ct, ci, vp]
        getattr(L, f"orc_apply_M_rows_{sfx}").argtypes = [vp, vp, ci, ct, ct, ci, vp, ci, ci, vp]
        getattr(L, f"orc_pair_block_{sfx}").argtypes = [vp, ci, ci, ct, ci, vp]
        getattr(L, f"orc_damp_{sfx}").argtypes = [vp, ci, ct, vp]
        getattr(L, f"orc_damp_{sfx}").restype = None
    L.orc_num_threads.restype = ci


class OracleError(RuntimeError):
    """Raised where the reference throws (blob below the wall) or exit()s (overlap)."""


def _check(err):
    if err == 1:
        raise OracleError("two blobs overlap (reference calls exit(), c_rigid_obj.cpp:53-58)")
    if err == 2:
        raise OracleError("A blob has its center below the wall (z<0) (c_rigid_obj.cpp:95-97)")
    if err:
        raise MemoryError("oracle: allocation failed")


def _prep(x, dtype):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype).reshape(-1))


# --------------------------------------------------------------------------- #
# O(N^2) mobility (ctypes into rbl_oracle.c)
# --------------------------------------------------------------------------- #
def dense_mobility(r, a, eta, wall, dtype=np.float64):
    """rotne_prager_tensor (c_rigid_obj.cpp:413-459): dense 3N x 3N matrix."""
    r = _prep(r, dtype)
    n = r.size // 3
    sfx, _ = _CT[np.dtype(dtype)]
    M = np.empty((3 * n, 3 * n), dtype=dtype, order="F")
    _check(getattr(lib(), f"orc_dense_mobility_{sfx}")(r.ctypes.data, n, a, eta, int(wall), M.ctypes.data))
    return M


def apply_M_dense(F, r, a, eta, wall, dtype=np.float64):
    """apply_M exactly as the reference runs it (c_rigid_obj.cpp:641-659):
    dense assembly + GEMV, single thread.  This is the timed CPU baseline."""
    F = _prep(F, dtype)
    r = _prep(r, dtype)
    n = r.size // 3
    sfx, _ = _CT[np.dtype(dtype)]
    U = np.empty(3 * n, dtype=dtype)
    _check(getattr(lib(), f"orc_apply_M_dense_{sfx}")(F.ctypes.data, r.ctypes.data, n, a, eta, int(wall), U.ctypes.data))
    return U


def apply_M(F, r, a, eta, wall, rows=None, dtype=np.float64):
    """Matrix-free rows of the same matrix (long-double row sums for float64).
    ``rows`` = None (all), an int array of target blob indices, or (row0, nrows)."""
    F = _prep(F, dtype)
    r = _prep(r, dtype)
    n = r.size // 3
    sfx, _ = _CT[np.dtype(dtype)]
    fn = getattr(lib(), f"orc_apply_M_rows_{sfx}")
    if rows is None:
        rows = (0, n)
    if isinstance(rows, tuple):
        row0, nrows = rows
        U = np.empty(3 * nrows, dtype=dtype)
        _check(fn(F.ctypes.data, r.ctypes.data, n, a, eta, int(wall), None, row0, nrows, U.ctypes.data))
    else:
        idx = np.ascontiguousarray(rows, dtype=np.int32)
        U = np.empty(3 * idx.size, dtype=dtype)
        _check(fn(F.ctypes.data, r.ctypes.data, n, a, eta, int(wall), idx.ctypes.data, 0, idx.size, U.ctypes.data))
    return U


def pair_block(r, i, j, a, wall, dtype=np.float64):
    """One upper-triangular block as rotne_prager_tensor's loop evaluates it (:432-447)."""
    r = _prep(r, dtype)
    sfx, _ = _CT[np.dtype(dtype)]
    B = np.empty(9, dtype=dtype)
    _check(getattr(lib(), f"orc_pair_block_{sfx}")(r.ctypes.data, i, j, a, int(wall), B.ctypes.data))
    return B.reshape(3, 3)


def damp_diag(r, a):
    """make_damp_mat (c_rigid_obj.cpp:618-639): per-blob B_ii = min(1, z/a) (3N)."""
    r = np.asarray(r, dtype=np.float64).reshape(-1, 3)
    d = np.where(r[:, 2] >= a, 1.0, r[:, 2] / a)
    return np.repeat(d, 3)


def num_threads():
    return int(lib().orc_num_threads())


# --------------------------------------------------------------------------- #
# O(N) rigid-body pieces (numpy float64)
# --------------------------------------------------------------------------- #
def remove_mean(cfg):
    """removeMean (c_rigid_obj.cpp:176-181)."""
    cfg = np.asarray(cfg, dtype=np.float64).reshape(-1, 3)
    return cfg - cfg.mean(axis=0)


def normalize_quats(Q):
    """setConfig normalises each quaternion (c_rigid_obj.cpp:212-216); layout [w,x,y,z]."""
    Q = np.asarray(Q, dtype=np.float64).reshape(-1, 4)
    return Q / np.linalg.norm(Q, axis=1, keepdims=True)


def rotation_matrices(Q):
    """Eigen Quaternion::toRotationMatrix for unit quaternions stored [w,x,y,z]
    (used at c_rigid_obj.cpp:258,308)."""
    Q = np.asarray(Q, dtype=np.float64).reshape(-1, 4)
    w, x, y, z = Q[:, 0], Q[:, 1], Q[:, 2], Q[:, 3]
    R = np.empty((Q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - w * z)
    R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y)
    R[:, 2, 1] = 2 * (y * z + w * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def blob_positions(X, Q, ref_cfg):
    """get_r_vecs ... multi_body_pos (c_rigid_obj.cpp:257-300):
    r_{b,k} = R(q_b) ref_k + X_b, body-major.  ``ref_cfg`` already mean-removed,
    ``Q`` already normalised.  Returns (N_bod*N_blb, 3)."""
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    R = rotation_matrices(Q)
    ref = np.asarray(ref_cfg, dtype=np.float64).reshape(-1, 3)
    r = np.einsum("bij,kj->bki", R, ref) + X[:, None, :]
    return r.reshape(-1, 3)


def _rho(r, X, n_blb):
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    return np.asarray(r, dtype=np.float64).reshape(X.shape[0], n_blb, 3) - X[:, None, :]


def K_dot(U, r, X, n_blb):
    """K_x_U (c_rigid_obj.cpp:404) with K's entries from :368-383:
    (K U)_k = u_b + omega_b x (r_k - X_b)."""
    U = np.asarray(U, dtype=np.float64).reshape(-1, 6)
    rho = _rho(r, X, n_blb)
    out = U[:, None, :3] + np.cross(U[:, None, 3:], rho)
    return out.reshape(-1)


def KT_dot(lam, r, X, n_blb):
    """KT_x_Lam (c_rigid_obj.cpp:410): F_b = sum_k lam_k, T_b = sum_k rho_k x lam_k."""
    rho = _rho(r, X, n_blb)
    lam = np.asarray(lam, dtype=np.float64).reshape(rho.shape)
    F = lam.sum(axis=1)
    T = np.cross(rho, lam).sum(axis=1)
    return np.concatenate([F, T], axis=1).reshape(-1)


def K_dense(r, X, n_blb):
    """Make_K_Kinv's K (c_rigid_obj.cpp:368-383) as a dense (3N, 6N_bod) array."""
    rho = _rho(r, X, n_blb)
    nb = rho.shape[0]
    K = np.zeros((3 * nb * n_blb, 6 * nb))
    for b in range(nb):
        for k in range(n_blb):
            row = 3 * (b * n_blb + k)
            rx, ry, rz = rho[b, k]
            K[row:row + 3, 6 * b:6 * b + 3] = np.eye(3)
            K[row + 0, 6 * b + 4] = rz
            K[row + 0, 6 * b + 5] = -ry
            K[row + 1, 6 * b + 5] = rx
            K[row + 1, 6 * b + 3] = -rz
            K[row + 2, 6 * b + 3] = ry
            K[row + 2, 6 * b + 4] = -rx
    return K


def KTK_inv_blocks(Q, ref_cfg):
    """block_KTKinv (c_rigid_obj.cpp:302-326): per body diag(I/N_blb, S),
    S = (sum|ref|^2 I - R (sum ref ref^T) R^T)^-1; off-diagonal blocks dropped."""
    ref = np.asarray(ref_cfg, dtype=np.float64).reshape(-1, 3)
    n_blb = ref.shape[0]
    sumr2 = (ref * ref).sum()
    moi = ref.T @ ref
    R = rotation_matrices(Q)
    out = np.zeros((R.shape[0], 6, 6))
    for b in range(R.shape[0]):
        D = sumr2 * np.eye(3) - R[b] @ moi @ R[b].T
        if np.linalg.det(D) < 1e-13:
            raise OracleError("K^T K is singular (reference exit()s, c_rigid_obj.cpp:313-316)")
        out[b, :3, :3] = np.eye(3) / n_blb
        out[b, 3:, 3:] = np.linalg.inv(D)
    return out


def Kinv_dense(r, X, Q, ref_cfg):
    """Kinv = (K^T K)^-1 K^T (c_rigid_obj.cpp:388-390), dense (6N_bod, 3N)."""
    ref = np.asarray(ref_cfg).reshape(-1, 3)
    K = K_dense(r, X, ref.shape[0])
    blk = KTK_inv_blocks(Q, ref)
    nb = blk.shape[0]
    G = np.zeros((6 * nb, 6 * nb))
    for b in range(nb):
        G[6 * b:6 * b + 6, 6 * b:6 * b + 6] = blk[b]
    return G @ K.T


def diag_invM(r, a, eta, wall):
    """diag_invM (c_rigid_obj.cpp:489-543): per blob inverse of the 3x3
    self-mobility (4/3 I + wall self term), times 8 pi eta a.  Returns (N,3,3)."""
    r = np.asarray(r, dtype=np.float64).reshape(-1, 3)
    n = r.shape[0]
    out = np.empty((n, 3, 3))
    for i in range(n):
        B = np.eye(3) * (4.0 / 3.0)
        if wall:
            h = r[i, 2] / a
            if h < 0:
                raise OracleError("A blob has its center below the wall (z<0)")
            inv = 1.0 / h
            i3 = inv ** 3
            i5 = i3 * inv * inv
            B[0, 0] += -(9 * inv - 2 * i3 + i5) / 12.0
            B[1, 1] += -(9 * inv - 2 * i3 + i5) / 12.0
            B[2, 2] += -(9 * inv - 4 * i3 + i5) / 6.0
        out[i] = np.linalg.inv(B) * (8.0 * np.pi * eta * a)
    return out


def block_invM(X, Q, ref_cfg, a, eta, wall):
    """Block_diag_invM (c_rigid_obj.cpp:461-487): per body, inverse of the dense
    RPY(+wall) matrix of that body's blobs alone.  Returns (N_bod, 3N_blb, 3N_blb)."""
    ref = np.asarray(ref_cfg, dtype=np.float64).reshape(-1, 3)
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    Q = np.asarray(Q, dtype=np.float64).reshape(-1, 4)
    nb = X.shape[0]
    sz = 3 * ref.shape[0]
    out = np.empty((nb, sz, sz))
    for b in range(nb):
        rb = blob_positions(X[b:b + 1], Q[b:b + 1], ref)
        out[b] = np.linalg.inv(np.asarray(dense_mobility(rb, a, eta, wall)))
    return out


class PC:
    """apply_PC (c_rigid_obj.cpp:589-616) with its lazily built state
    (:591-596): invM (diag :489-543 or block :461-487), Ninv = K^T invM K per body
    (:593) and its Cholesky solve (:554-567, :605-608).  M_scale = 1 (:194)."""

    def __init__(self, X, Q, ref_cfg, a, eta, wall, block):
        self.ref = np.asarray(ref_cfg, dtype=np.float64).reshape(-1, 3)
        self.n_blb = self.ref.shape[0]
        self.X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
        self.Q = np.asarray(Q, dtype=np.float64).reshape(-1, 4)
        self.nb = self.X.shape[0]
        self.r = blob_positions(self.X, self.Q, self.ref)
        sz = 3 * self.n_blb
        if block:
            self.invM = block_invM(self.X, self.Q, self.ref, a, eta, wall)
        else:
            d = diag_invM(self.r, a, eta, wall)
            self.invM = np.zeros((self.nb, sz, sz))
            for b in range(self.nb):
                for k in range(self.n_blb):
                    self.invM[b, 3 * k:3 * k + 3, 3 * k:3 * k + 3] = d[b * self.n_blb + k]
        Kd = K_dense(self.r, self.X, self.n_blb)
        self.Kb = [Kd[sz * b:sz * (b + 1), 6 * b:6 * b + 6] for b in range(self.nb)]
        self.Ninv = [self.Kb[b].T @ self.invM[b] @ self.Kb[b] for b in range(self.nb)]

    def apply(self, IN):
        IN = np.asarray(IN, dtype=np.float64).reshape(-1)
        sz = 3 * self.n_blb
        n3 = sz * self.nb
        slip = IN[:n3]
        F = IN[n3:]
        lam = np.empty(n3)
        U = np.empty(6 * self.nb)
        for b in range(self.nb):
            s = slip[sz * b:sz * (b + 1)]
            rhs = -F[6 * b:6 * b + 6] - self.Kb[b].T @ (self.invM[b] @ s)  # :601
            L = np.linalg.cholesky(self.Ninv[b])                            # :562
            u = np.linalg.solve(L.T, np.linalg.solve(L, rhs))               # :607
            U[6 * b:6 * b + 6] = u
            lam[sz * b:sz * (b + 1)] = self.invM[b] @ (s + self.Kb[b] @ u)  # :610
        return np.concatenate([lam, U])


def apply_saddle(x, X, Q, ref_cfg, a, eta, wall):
    """RigidBody.apply_saddle (Rigid.py:73-80): [M lam - K U ; K^T lam]."""
    ref = np.asarray(ref_cfg, dtype=np.float64).reshape(-1, 3)
    n_blb = ref.shape[0]
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    r = blob_positions(X, Q, ref)
    n3 = 3 * r.shape[0]
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    lam, U = x[:n3], x[n3:]
    slip = apply_M(lam, r, a, eta, wall) - K_dot(U, r, X, n_blb)
    return np.concatenate([slip, KT_dot(lam, r, X, n_blb)])


def quat_mul(p, q):
    """Hamilton product p*q for arrays stored [w,x,y,z] (Eigen operator*, :704)."""
    pw, px, py, pz = p[..., 0], p[..., 1], p[..., 2], p[..., 3]
    qw, qx, qy, qz = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([
        pw * qw - px * qx - py * qy - pz * qz,
        pw * qx + px * qw + py * qz - pz * qy,
        pw * qy - px * qz + py * qw + pz * qx,
        pw * qz + px * qy - py * qx + pz * qw,
    ], axis=-1)


def update_X_Q(X, Q, disp):
    """update_X_Q + Q_from_Om (c_rigid_obj.cpp:679-710): ``disp`` already has
    units of displacement ([u(3), omega(3)] * dt per body)."""
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    Q = np.asarray(Q, dtype=np.float64).reshape(-1, 4)
    d = np.asarray(disp, dtype=np.float64).reshape(-1, 6)
    om = d[:, 3:]
    th = np.linalg.norm(om, axis=1)
    qrot = np.zeros_like(Q)
    qrot[:, 0] = np.cos(th / 2)
    big = th > 1e-10
    qrot[big, 1:] = (np.sin(th[big] / 2) / th[big])[:, None] * om[big]
    qrot /= np.linalg.norm(qrot, axis=1, keepdims=True)
    Qn = quat_mul(qrot, Q)
    Qn /= np.linalg.norm(Qn, axis=1, keepdims=True)
    return X + d[:, :3], Qn


def evolve(X, Q, U, dt):
    """evolve_X_Q (c_rigid_obj.cpp:865-878): U *= dt, then update_X_Q."""
    return update_X_Q(X, Q, np.asarray(U, dtype=np.float64) * dt)


def Kinv_apply(v, r, X, Q, ref_cfg):
    """Kinv_x_V (c_rigid_obj.cpp:406): (K^T K)^-1 K^T v with the closed-form blocks (:302-326)."""
    ref = np.asarray(ref_cfg, dtype=np.float64).reshape(-1, 3)
    y = KT_dot(v, r, X, ref.shape[0]).reshape(-1, 6)
    G = KTK_inv_blocks(Q, ref)
    return np.einsum("bij,bj->bi", G, y).reshape(-1)


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11; Random123) on arrays of counters: ``counter`` = four
    uint32 arrays, ``key`` = two uint32 scalars.  Returns four uint32 arrays.  Checked against the
    Random123 known-answer vectors in tests/test_oracle_physics.py."""
    c = [np.asarray(x, dtype=np.uint64) for x in counter]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1, mask, s32 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF), np.uint64(32)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> s32) ^ c[1] ^ k0, p1 & mask, (p0 >> s32) ^ c[3] ^ k1, p0 & mask]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c


def philox_normals(seed, step, first, n):
    """The three standard-normal vectors (W1, W2, Wr) of BD step ``step`` for global elements
    [first, first + n): element e uses ONE Philox block with counter (e_lo, e_hi, step_lo, step_hi)
    and key (seed_lo, seed_hi); W1, W2 = Box-Muller pair of words 0,1; Wr = cosine branch of words
    2,3 (rbl_krylov.cu normal_triplet_kernel; include/rbl.h rbl_bd_step_seeded)."""
    e = np.uint64(first) + np.arange(n, dtype=np.uint64)
    m, s32 = np.uint64(0xFFFFFFFF), np.uint64(32)
    step, seed = np.uint64(step), np.uint64(seed)
    r = philox4x32_10([e & m, e >> s32, np.full(n, step & m), np.full(n, step >> s32)], [seed & m, seed >> s32])
    u = [(x.astype(np.float64) + 0.5) * 2.3283064365386963e-10 for x in r]
    ra, rb = np.sqrt(-2.0 * np.log(u[0])), np.sqrt(-2.0 * np.log(u[2]))
    return ra * np.cos(2 * np.pi * u[1]), ra * np.sin(2 * np.pi * u[1]), rb * np.cos(2 * np.pi * u[3])


def noise_factors(r, Q, ref_cfg, a, eta, wall):
    """Per-body factors L_b with L_b L_b^T = Mt_b, the body's own mobility block (no B damping),
    as the product path builds them: with the wall, the lower Cholesky factor of each body's block;
    in free space ONE Cholesky factor of the reference shape, rotated -- M_b = R M_ref R^T with
    R = blockdiag(R_b) per blob, so L_b = R L_ref (not triangular, equally valid)."""
    ref = np.asarray(ref_cfg, dtype=np.float64).reshape(-1, 3)
    n_blb = ref.shape[0]
    r = np.asarray(r, dtype=np.float64).reshape(-1, 3)
    nb = r.shape[0] // n_blb
    out = []
    if wall:
        for b in range(nb):
            Mb = np.asarray(dense_mobility(r[b * n_blb:(b + 1) * n_blb], a, eta, True))
            out.append(np.linalg.cholesky(Mb))
    else:
        L_ref = np.linalg.cholesky(np.asarray(dense_mobility(ref, a, eta, False)))
        R = rotation_matrices(normalize_quats(np.asarray(Q, dtype=np.float64).reshape(-1, 4)))
        for b in range(nb):
            out.append(np.kron(np.eye(n_blb), R[b]) @ L_ref)
    return out


def noise_block_cholesky(factors, A, W):
    """g = L (G A G^T)^{1/2} W with L = blockdiag(factors) and G = L^-1: a vector of covariance A
    (= B M B) through the block preconditioned square root (include/rbl.h
    rbl_set_noise_preconditioner).  Dense float64."""
    from scipy.linalg import sqrtm

    n = A.shape[0]
    L = np.zeros((n, n))
    o = 0
    for Lb in factors:
        sz = Lb.shape[0]
        L[o:o + sz, o:o + sz] = Lb
        o += sz
    G = np.linalg.inv(L)
    S = np.real(sqrtm(G @ A @ G.T))
    return L @ (S @ W)


def rfd_M(X, Q, ref_cfg, a, eta, wall, W, delta=1.0e-4):
    """M_RFD (c_rigid_obj.cpp:769-796): (M(q+) - M(q-)) W / delta with q+- = q +- (delta/2) K^-1 W."""
    ref = np.asarray(ref_cfg, dtype=np.float64).reshape(-1, 3)
    r = blob_positions(X, Q, ref)
    uom = Kinv_apply(W, r, X, Q, ref)
    Xp, Qp = update_X_Q(X, Q, 0.5 * delta * uom)
    Xn, Qn = update_X_Q(X, Q, -0.5 * delta * uom)
    rp, rn = blob_positions(Xp, Qp, ref), blob_positions(Xn, Qn, ref)
    return (apply_M(W, rp, a, eta, wall) - apply_M(W, rn, a, eta, wall)) / delta


def bd_step(X, Q, ref_cfg, a, eta, dt, kBT, wall, F_ext, slip, W1, W2, Wr, delta=1.0e-4, noise="symmetric",
            return_rhs=False, split_rand=True):
    """The trapezoidal-slip midpoint step RHS_and_Midpoint sets up (c_rigid_obj.cpp:917-976),
    completed as intended (the reference computes the midpoint configuration but never installs
    it) and evaluated with dense float64 linear algebra:
      M^{1/2}W by scipy sqrtm of B M B (the reference uses the Cholesky factor, :661-675 -- a
      different square root with the same covariance; Lanczos converges to the symmetric one),
      RFD (:769-796), BI (:945-948), midpoint (:954-958), dense solve of
      [M -K; K^T 0][lam;U] = [slip - kBT RFD - BI ; F_ext] at the midpoint, evolve from q^n.
    Returns (U, X_new, Q_new); with return_rhs the assembled right-hand side and the midpoint
    configuration instead (rhs, X_mid, Q_mid).  noise: "symmetric" (sqrtm), "block_cholesky" (the
    product path's default) or "cholesky" (the reference's M_half_W).  split_rand=False: the single-increment
    branch (:949-953): c1 = c2 = sqrt(2 kBT/dt), BI = c2 M^{1/2}W1, W2 unused."""
    from scipy.linalg import sqrtm

    if W2 is None:
        W2 = np.zeros_like(np.asarray(W1, dtype=np.float64))
    ref = np.asarray(ref_cfg, dtype=np.float64).reshape(-1, 3)
    n_blb = ref.shape[0]
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    Q = np.asarray(Q, dtype=np.float64).reshape(-1, 4)
    nb = X.shape[0]
    r = blob_positions(X, Q, ref)
    n3 = r.size
    rhs_slip = np.zeros(n3) if slip is None else np.asarray(slip, dtype=np.float64).reshape(-1).copy()
    Xm, Qm = X, Q
    if kBT > 0:
        M_raw = np.asarray(dense_mobility(r, a, eta, wall))
        M = M_raw
        if wall:
            B = damp_diag(r, a)
            M = B[:, None] * M_raw * B[None, :]
        if noise == "block_cholesky":
            fac = noise_factors(r, Q, ref, a, eta, wall)
            mh1 = noise_block_cholesky(fac, M, W1)
            mh2 = noise_block_cholesky(fac, M, W2)
        elif noise == "cholesky":  # the reference's own M_half_W (:661-675): the lower Cholesky factor of B M B
            Lc = np.linalg.cholesky(M)
            mh1, mh2 = Lc @ W1, Lc @ W2
        else:
            S = np.real(sqrtm(M))
            mh1, mh2 = S @ W1, S @ W2
        rfd = rfd_M(X, Q, ref, a, eta, wall, Wr, delta)
        if split_rand:
            c1, c2 = 2.0 * np.sqrt(kBT / dt), np.sqrt(kBT / dt)
            rhs_slip -= kBT * rfd + c2 * (mh1 - mh2)
        else:
            c1 = c2 = np.sqrt(2.0 * kBT / dt)
            rhs_slip -= kBT * rfd + c2 * mh1
        Xm, Qm = update_X_Q(X, Q, 0.5 * dt * Kinv_apply(c1 * mh1, r, X, Q, ref))
    rm = blob_positions(Xm, Qm, ref)
    Mm = np.asarray(dense_mobility(rm, a, eta, wall))
    if wall:
        B = damp_diag(rm, a)
        Mm = B[:, None] * Mm * B[None, :]
    K = K_dense(rm, Xm, n_blb)
    A = np.block([[Mm, -K], [K.T, np.zeros((6 * nb, 6 * nb))]])
    rhs = np.concatenate([rhs_slip, np.asarray(F_ext, dtype=np.float64).reshape(-1)])
    if return_rhs:
        return rhs, Xm, Qm
    sol = np.linalg.solve(A, rhs)
    U = sol[n3:]
    Xn, Qn = evolve(X, Q, U, dt)
    return U, Xn, Qn
