"""Small end-to-end pass over every kernel for compute-sanitizer (developer tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from Rigid import RigidBody
from rigid_body_light_b200.shells import sphere_suspension
from rigid_body_light_b200._lib import Context
rng = np.random.default_rng(0)
for precision in ("single", "double"):
    for wall in (True, False):
        for block in (False, True):
            s = sphere_suspension(7, 42, wall)
            cb = RigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=wall, block_PC=block, precision=precision)
            n3, n6 = 3 * 7 * 42, 42
            v = rng.standard_normal(n3 + n6)
            cb.apply_saddle(v); cb.apply_PC(v); cb.K_dot(v[n3:]); cb.KT_dot(v[:n3]); cb.Kinv_dot(v[:n3]); cb.KTinv_dot(v[n3:])
            cb.get_K(); cb.get_Kinv()
            cb.gmres(v, tol=1e-4, restart=20, max_iter=40)
            cb.brownian_sqrt(v[:n3], tol=1e-3, max_iter=20)
            cb.bd_step(v[n3:], kBT=0.01, rng=rng, tol=1e-4, restart=20, max_iter=40, lanczos_tol=1e-3, lanczos_max_iter=20)
    # ragged product sizes through both kernels and every variant
    ctx = Context(precision)
    ctx.set_parameters(0.1, 0.01, 1.0, 1.0, np.zeros((1, 3)))
    for wall in (0, 1):
        ctx.set_flags(0, wall)
        for n in (1, 33, 257, 1025, 2300):
            r = rng.uniform(0.3, 8, (n, 3)); F = rng.standard_normal(3 * n)
            for mode in (0, 1):
                ctx.call("rbl_set_matvec_mode", mode)
                nv = ctx.L.rbl_num_sym_variants(ctx.h) if mode == 0 else ctx.L.rbl_num_matvec_variants(ctx.h)
                for vv in range(nv):
                    ctx.call("rbl_set_sym_variant" if mode == 0 else "rbl_set_matvec_variant", vv)
                    ctx.apply_M(F, r)
    ctx.close()
print("sanitize pass done")
