"""``RigidBody`` -- the Python API of Rigid_Body_Light on the B200-native host class.

Mirror of /root/reference/src/Rigid.py:5-135: same constructor signature, same method
names, argument meaning, return shapes and ``RuntimeError`` conditions, so code and tests
written against the reference run unchanged.  Differences, all additive:

* ``precision=`` keyword (default: the module default, like the reference's compile-time
  choice) selects the float or double host class per object;
* ``apply_saddle`` calls the fused device operator instead of composing four host calls
  (same result: [M lam - K U ; K^T lam], Rigid.py:73-80);
* ``Kinv_dot`` / ``KTinv_dot`` / ``gmres`` / ``brownian_sqrt`` expose operators the
  reference keeps unbound or leaves to the user.
"""
import numpy as np

from . import c_rigid as crigid


def _count(x):
    return int(np.prod(np.shape(x)))


class RigidBody:
    X_shape = None
    Q_shape = None

    def __init__(self, rigid_config, X, Q, a, eta, dt, wall_PC=False, block_PC=False, precision=None):
        rigid_config = np.asarray(rigid_config)
        if rigid_config.size % 3 != 0:
            raise RuntimeError(
                f"Rigid config must have length 3N. Rigid config shape: {rigid_config.shape}"
            )
        self.cb = crigid.host_class(precision)()
        self.precision = self.cb.precision
        self.blobs_per_body = rigid_config.size // 3
        kbt = 1.0  # the reference hard-wires kBT = 1 here too (Rigid.py:23)
        self.cb.setParameters(a, dt, kbt, eta, rigid_config)
        self.cb.setBlkPC(block_PC)
        self.cb.setWallPC(wall_PC)
        self.set_config(X, Q)

    # -- configuration ---------------------------------------------------------------
    def set_config(self, X, Q):
        X = np.asarray(X)
        Q = np.asarray(Q)
        nx, rx = divmod(_count(X), 3)
        nq, rq = divmod(_count(Q), 4)
        if rx:
            raise RuntimeError("X must have total length 3N")
        if rq:
            raise RuntimeError("Q must have total length 4N")
        if nx != nq:
            raise RuntimeError("X and Q must have the same number of bodies")
        self.N_bodies = nx
        self.X_shape = X.shape
        self.Q_shape = Q.shape
        self.cb.setConfig(X.reshape(-1), Q.reshape(-1))
        self.cb.set_K_mats()
        self.total_blobs = self.N_bodies * self.blobs_per_body

    def get_config(self):
        X, Q = self.cb.getConfig()
        return X.reshape(self.X_shape), Q.reshape(self.Q_shape)

    def _like_X(self, flat):
        return np.asarray(flat).reshape((-1, 3) if len(self.X_shape) == 2 else (-1))

    def get_blob_positions(self):
        return self._like_X(self.cb.multi_body_pos())

    # -- operators -------------------------------------------------------------------
    def KT_dot(self, lambda_vec):
        lambda_vec = np.asarray(lambda_vec)
        self._need(lambda_vec, 3 * self.total_blobs, "lambda", "3*N_blobs")
        return self._like_X(self.cb.KT_x_Lam(lambda_vec.reshape(-1)))

    def K_dot(self, U):
        U = np.asarray(U)
        self._need(U, 6 * self.N_bodies, "U", "6*N_bodies")
        return self._like_X(self.cb.K_x_U(U.reshape(-1)))

    def apply_PC(self, b):
        b = np.asarray(b)
        self._need_system(b)
        return self.cb.apply_PC(b.reshape(-1))

    def apply_saddle(self, x):
        x = np.asarray(x)
        self._need_system(x)
        return self.cb.apply_saddle(x.reshape(-1))

    def apply_M(self, forces, positions):
        forces = np.asarray(forces)
        positions = np.asarray(positions)
        if np.size(positions) != np.size(forces):
            raise RuntimeError("Positions and forces must be of the same size")
        if np.size(positions) % 3 != 0:
            raise RuntimeError(
                "Positions and forces must have total length 3N, where N is the number of blobs"
            )
        return self.cb.apply_M(forces.reshape(-1), positions.reshape(-1))

    def get_K(self):
        return self.cb.get_K()

    def get_Kinv(self):
        return self.cb.get_Kinv()

    def evolve_rigid_bodies(self, U):
        U = np.asarray(U)
        self._need(U, 6 * self.N_bodies, "U", "6*N_bodies")
        self.cb.evolve_X_Q(U.reshape(-1))

    # -- extensions (not in the reference's Python API) ---------------------------------
    def Kinv_dot(self, V):
        V = np.asarray(V)
        self._need(V, 3 * self.total_blobs, "V", "3*N_blobs")
        return self._like_X(self.cb.Kinv_x_V(V.reshape(-1)))

    def KTinv_dot(self, F):
        F = np.asarray(F)
        self._need(F, 6 * self.N_bodies, "F", "6*N_bodies")
        return self._like_X(self.cb.KTinv_x_F(F.reshape(-1)))

    def gmres(self, rhs, tol=1e-8, restart=60, max_iter=300):
        """Solve apply_saddle(x) = rhs on the device, preconditioned with apply_PC.
        Returns (x, iterations, relative residual)."""
        rhs = np.asarray(rhs)
        self._need_system(rhs)
        return self.cb.gmres(rhs.reshape(-1), tol, restart, max_iter)

    def brownian_sqrt(self, W, tol=1e-6, max_iter=100):
        """(B M B)^{1/2} W by Lanczos over the device mobility product.
        Returns (vector, iterations)."""
        W = np.asarray(W)
        self._need(W, 3 * self.total_blobs, "W", "3*N_blobs")
        return self.cb.lanczos_sqrt(W.reshape(-1), tol, max_iter)

    def apply_M2(self, forces1, forces2, positions):
        """(M F1, M F2) in one pass over the blob pairs (two-right-hand-side kernel)."""
        f1, f2, r = np.asarray(forces1), np.asarray(forces2), np.asarray(positions)
        if f1.size != r.size or f2.size != r.size:
            raise RuntimeError("Positions and forces must be of the same size")
        if r.size % 3 != 0:
            raise RuntimeError("Positions and forces must have total length 3N, where N is the number of blobs")
        return self.cb.apply_M2(f1.reshape(-1), f2.reshape(-1), r.reshape(-1))

    def brownian_sqrt_pair(self, W1, W2, tol=1e-6, max_iter=100):
        """(B M B)^{1/2} W1 and (B M B)^{1/2} W2 by two Lanczos recurrences that share every
        mobility product.  Returns (vector1, vector2, iterations1, iterations2)."""
        W1, W2 = np.asarray(W1), np.asarray(W2)
        self._need(W1, 3 * self.total_blobs, "W1", "3*N_blobs")
        self._need(W2, 3 * self.total_blobs, "W2", "3*N_blobs")
        return self.cb.lanczos_sqrt2(W1.reshape(-1), W2.reshape(-1), tol, max_iter)

    def set_noise_preconditioner(self, mode):
        """0: Brownian increments through the symmetric square root (B M B)^{1/2} W; 1 (default):
        block-Cholesky preconditioned increments L (G B M B G^T)^{1/2} W inside ``bd_step`` (same
        covariance, ~3x fewer mobility products); 2: ``brownian_sqrt`` / ``brownian_sqrt_pair``
        return the preconditioned vector too."""
        self.cb.set_noise_preconditioner(int(mode))

    def bd_step(self, F_ext, slip=None, kBT=0.0, noise=None, rng=None, tol=1e-8, restart=60, max_iter=300,
                lanczos_tol=1e-6, lanczos_max_iter=100, seed=None, step=0):
        """Advance the bodies by one (Brownian) step with the trapezoidal-slip midpoint scheme
        the reference sets up in RHS_and_Midpoint (c_rigid_obj.cpp:917-976), entirely on the
        device.  ``F_ext``: external force/torque per body (6*N_bodies); ``slip``: prescribed
        blob slip (3*N_blobs) or None; ``noise`` = (W1, W2, Wr) standard-normal 3*N_blobs
        vectors, drawn from ``rng`` (numpy Generator) when kBT > 0 and noise is None; or pass
        ``seed`` (and the ``step`` number) to draw them on the device with the counter-based generator
        (a pure function of seed, step and blob index: reproducible, partition invariant).
        Returns (U, gmres_iterations, relative_residual)."""
        F_ext = np.asarray(F_ext)
        self._need(F_ext, 6 * self.N_bodies, "F_ext", "6*N_bodies")
        n3 = 3 * self.total_blobs
        if slip is not None:
            slip = np.asarray(slip)
            self._need(slip, n3, "slip", "3*N_blobs")
            slip = slip.reshape(-1)
        if seed is not None and kBT > 0 and noise is None:
            return self.cb.bd_step_seeded(F_ext.reshape(-1), slip, int(seed), int(step), float(kBT), tol, restart, max_iter,
                                          lanczos_tol, lanczos_max_iter)
        W1 = W2 = Wr = None
        if kBT > 0:
            if noise is None:
                rng = np.random.default_rng() if rng is None else rng
                noise = tuple(rng.standard_normal(n3) for _ in range(3))
            W1, W2, Wr = (np.asarray(w).reshape(-1) for w in noise)
        return self.cb.bd_step(F_ext.reshape(-1), slip, W1, W2, Wr, float(kBT), tol, restart, max_iter,
                               lanczos_tol, lanczos_max_iter)

    # -- random finite differences: the reference's unbound members (c_rigid_obj.cpp:712-728,743-863,
    #    880-893) through the C ABI (include/rbl.h), noise supplied by the caller -----------------
    def _abi(self, name, *args):
        import ctypes

        from . import _lib

        L = _lib.load()
        st = getattr(L, name)(ctypes.c_void_p(self.cb.handle()), *args)
        if st != 0:
            raise RuntimeError(L.rbl_last_error(ctypes.c_void_p(self.cb.handle())).decode())

    def _real(self):
        return np.float64 if self.precision == "double" else np.float32

    def _vec(self, x, n, name, what):
        x = np.asarray(x)
        self._need(x, n, name, what)
        return np.ascontiguousarray(x.reshape(-1), dtype=self._real())

    def M_RFD(self, W, delta=0.0, U=None):
        """(M(q+) - M(q-)) W / delta with q+- = q +- (delta/2) K^-1 W (M_RFD, :769-796), or with
        q+- = q +- (delta/2) U when ``U`` is given (M_RFD_from_U, :818-840).  delta <= 0: the reference's."""
        W = self._vec(W, 3 * self.total_blobs, "W", "3*N_blobs")
        out = np.empty_like(W)
        if U is None:
            self._abi("rbl_M_RFD", W.ctypes.data, float(delta), out.ctypes.data)
        else:
            U = self._vec(U, 6 * self.N_bodies, "U", "6*N_bodies")
            self._abi("rbl_M_RFD_from_U", U.ctypes.data, W.ctypes.data, float(delta), out.ctypes.data)
        return out

    def KT_RFD(self, U, W, delta=0.0):
        """(K(q+)^T - K(q-)^T) W / delta, q+- = q +- (delta/2) U (KT_RFD_from_U, :842-863)."""
        U = self._vec(U, 6 * self.N_bodies, "U", "6*N_bodies")
        W = self._vec(W, 3 * self.total_blobs, "W", "3*N_blobs")
        out = np.empty(6 * self.N_bodies, dtype=self._real())
        self._abi("rbl_KT_RFD_from_U", U.ctypes.data, W.ctypes.data, float(delta), out.ctypes.data)
        return out

    def KTinv_RFD(self, W6, delta=0.0):
        """K^T (Kinv(q+)^T - Kinv(q-)^T) W / delta, q+- = q +- (delta/2) W (KTinv_RFD, :743-767)."""
        W6 = self._vec(W6, 6 * self.N_bodies, "W", "6*N_bodies")
        out = np.empty_like(W6)
        self._abi("rbl_KTinv_RFD", W6.ctypes.data, float(delta), out.ctypes.data)
        return out

    def M_RFD_cfgs(self, U, delta):
        """Blob positions of q +- (delta/2) U (M_RFD_cfgs, :798-816): (r_plus, r_minus), each (N_blobs, 3)."""
        U = self._vec(U, 6 * self.N_bodies, "U", "6*N_bodies")
        rp = np.empty(3 * self.total_blobs, dtype=self._real())
        rm = np.empty_like(rp)
        self._abi("rbl_M_RFD_cfgs", U.ctypes.data, float(delta), rp.ctypes.data, rm.ctypes.data)
        return rp.reshape(-1, 3), rm.reshape(-1, 3)

    def displaced_config(self, U):
        """(X, Q) displaced by the displacement U, state untouched (update_X_Q_out, :712-728)."""
        U = self._vec(U, 6 * self.N_bodies, "U", "6*N_bodies")
        X = np.empty(3 * self.N_bodies, dtype=self._real())
        Q = np.empty(4 * self.N_bodies, dtype=self._real())
        self._abi("rbl_update_X_Q_out", U.ctypes.data, X.ctypes.data, Q.ctypes.data)
        return X.reshape(-1, 3), Q.reshape(-1, 4)

    def evolve_rigid_bodies_RFD(self, U):
        """Install the configuration displaced by U (a displacement) and rebuild K; a built preconditioner
        is kept (evolve_X_Q_RFD, :880-893)."""
        U = self._vec(U, 6 * self.N_bodies, "U", "6*N_bodies")
        self._abi("rbl_evolve_RFD", U.ctypes.data)

    def set_rfd_delta(self, delta):
        """RFD step of ``bd_step`` / default of ``M_RFD``; 0 restores 1e-4 (double) / 4e-3 (float)."""
        self._abi("rbl_set_rfd_delta", float(delta))

    def set_mixed_precision(self, mode):
        """double contexts: 0 all double; 1 float GMRES corrections inside a double-residual refinement;
        2 also float mobility products inside the Lanczos noise (include/rbl.h rbl_set_mixed_precision)."""
        self._abi("rbl_set_mixed_precision", int(mode))

    def set_split_rand(self, enable):
        """True (default): two Brownian increments per step (:943-948); False: one (:949-953)."""
        self._abi("rbl_set_split_rand", int(bool(enable)))

    # -- size checks (RuntimeError like Rigid.py:117-135) -------------------------------
    def _need(self, vec, n, name, what):
        if vec.size != n:
            raise RuntimeError(
                f"{name} must have total size {what} = {n}. {name} shape: {vec.shape}"
            )

    def _need_system(self, vec):
        n = 3 * self.total_blobs + 6 * self.N_bodies
        if vec.size != n:
            raise RuntimeError(
                "Rigid system input vector must have total size 3*N_blobs + 6*N_bodies = "
                f"{n}. system_input shape: {vec.shape}"
            )
