"""Runs the BASELINE.json configurations that are not the bench line and prints one JSON
line each (GPU box).  cfg1: deterministic GMRES mobility solve (double) with the CPU path the
reference offers its users (scipy gmres over the oracle's operators) timed beside it;
cfg3: one full fluctuating BD step (GMRES + 2 Lanczos + RFD) with the wall."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from Rigid import RigidBody  # noqa: E402
from rigid_body_light_b200.shells import sphere_suspension  # noqa: E402


def cfg1(args):
    s = sphere_suspension(100, 42, True)
    n3, n6 = 3 * 4200, 600
    rng = np.random.default_rng(2)
    rhs = np.concatenate([np.zeros(n3), rng.standard_normal(n6)])
    out = {"config": "cfg1: 100 spheres of shell_N_42 above a wall, one deterministic GMRES mobility solve, double"}
    for block in (True, False):
        cb = RigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, block_PC=block, precision="double")
        cb.gmres(rhs, tol=1e-8, restart=60, max_iter=300)  # warm-up (PC build, allocations)
        t0 = time.perf_counter()
        x, iters, relres = cb.gmres(rhs, tol=1e-8, restart=60, max_iter=300)
        dt = time.perf_counter() - t0
        res = np.linalg.norm(cb.apply_saddle(x) - rhs) / np.linalg.norm(rhs)
        out["gpu_block_pc" if block else "gpu_diag_pc"] = {"seconds": dt, "iterations": int(iters), "relres": float(relres),
                                                             "true_residual": float(res)}
    cbm = RigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, block_PC=True, precision="double")
    cbm.set_mixed_precision(1)  # float GMRES corrections, double residual (same stopping rule)
    cbm.gmres(rhs, tol=1e-8, restart=60, max_iter=300)
    t0 = time.perf_counter()
    xm, iters, relres = cbm.gmres(rhs, tol=1e-8, restart=60, max_iter=300)
    dt = time.perf_counter() - t0
    out["gpu_block_pc_mixed_precision"] = {"seconds": dt, "float_iterations": int(iters), "relres": float(relres),
                                           "true_residual": float(np.linalg.norm(cb.apply_saddle(xm) - rhs) / np.linalg.norm(rhs))}
    if args.cpu:
        from scipy.sparse.linalg import LinearOperator, gmres

        from oracle import oracle as orc

        ref = orc.remove_mean(s["cfg"])
        n = n3 + n6
        A = LinearOperator((n, n), matvec=lambda v: orc.apply_saddle(v, s["X"], s["Q"], ref, s["a"], 1.0, True))
        t0 = time.perf_counter()
        pc = orc.PC(s["X"], s["Q"], ref, s["a"], 1.0, True, True)
        t_pc = time.perf_counter() - t0
        flip = np.concatenate([np.ones(n3), -np.ones(n6)])
        P = LinearOperator((n, n), matvec=lambda v: pc.apply(flip * v))
        count = [0]
        t0 = time.perf_counter()
        xc, info = gmres(A, rhs, M=P, rtol=1e-8, restart=60, maxiter=300, callback=lambda r: count.__setitem__(0, count[0] + 1),
                         callback_type="pr_norm")
        dt = time.perf_counter() - t0
        out["cpu_scipy_gmres_oracle_ops"] = {"seconds": dt, "pc_build_seconds": t_pc, "iterations": count[0], "info": int(info),
                                             "threads": orc.num_threads(),
                                             "vs_gpu_solution": float(np.linalg.norm(xc - x) / np.linalg.norm(x))}
        if orc.ref_apply_M_lib() is not None:
            # THE REFERENCE PATH ITSELF: scipy gmres over the reference's own members (compiled from its
            # source, oracle.RefBody) composed exactly like src/Rigid.py:73-80 -- every operator application
            # re-places the blobs, assembles the dense 12600 x 12600 matrix and runs a GEMV, single thread.
            # Block PC like the fastest GPU run (the shim stores the reference's sparse matrices densely, so
            # the PC set-up time is reported apart: Eigen's sparse products would be faster there).
            rb = orc.RefBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, block_PC=True)

            def ref_saddle(v):
                lam, U = v[:n3], v[n3:]
                r = rb.positions()
                return np.concatenate([rb.apply_M(lam, r) - rb.K_dot(U), rb.KT_dot(lam)])

            t0 = time.perf_counter()
            rb.apply_PC(np.zeros(n))
            t_pc = time.perf_counter() - t0
            Pr = LinearOperator((n, n), matvec=lambda v: rb.apply_PC(flip * v))
            Ar = LinearOperator((n, n), matvec=ref_saddle)
            count = [0]
            t0 = time.perf_counter()
            xr, info = gmres(Ar, rhs, M=Pr, rtol=1e-8, restart=60, maxiter=300, callback=lambda r_: count.__setitem__(0, count[0] + 1),
                             callback_type="pr_norm")
            dt = time.perf_counter() - t0
            out["cpu_reference_members_scipy_gmres"] = {
                "seconds": dt, "pc_build_seconds_shim_dense": t_pc, "iterations": count[0], "info": int(info), "threads": 1,
                "preconditioner": "block", "vs_gpu_solution": float(np.linalg.norm(xr - x) / np.linalg.norm(x))}
    print(json.dumps(out), flush=True)


def cfg3(args):
    s = sphere_suspension(4096, 42, True)
    nb = 4096
    F = np.tile(np.array([0, 0, -1.0, 0, 0, 0]), nb)  # gravity
    out = {"config": "cfg3: 4096 spheres of shell_N_42 (172 032 blobs), full fluctuating BD timestep with wall"}
    for precision in ("single", "double"):
        cb = RigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, block_PC=False, precision=precision)
        rng = np.random.default_rng(3)
        tol = 1e-4 if precision == "single" else 1e-8
        ltol = 1e-4 if precision == "single" else 1e-6
        t0 = time.perf_counter()
        U, iters, relres = cb.bd_step(F, kBT=0.0041, rng=rng, tol=tol, restart=60, max_iter=200, lanczos_tol=ltol,
                                      lanczos_max_iter=60)
        dt = time.perf_counter() - t0
        X, Q = cb.get_config()
        out[precision] = {"seconds_per_bd_step": dt, "gmres_iterations": int(iters), "relres": float(relres),
                          "min_z_after": float(X[:, 2].min()), "U_norm": float(np.linalg.norm(U))}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="cfg1,cfg3")
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    for w in a.which.split(","):
        {"cfg1": cfg1, "cfg3": cfg3}[w](a)
