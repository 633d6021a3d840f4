import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_light_b200._lib import Context
for precision in ("single", "double"):
    for n in (300, 2688, 4099, 20000):
        for wall in (1, 0):
            rng = np.random.default_rng(n)
            r = rng.uniform(0.5, 30, (n, 3)); F = rng.standard_normal(3 * n)
            ctx = Context(precision)
            ctx.set_parameters(0.1, 0.01, 1.0, 1.0, np.zeros((1, 3))); ctx.set_flags(0, wall)
            nv = ctx.L.rbl_num_matvec_variants(ctx.h)
            for v in [-1] + list(range(nv)):
                ctx.call("rbl_set_matvec_variant", v)
                ref = ctx.apply_M(F, r); bad = 0
                for k in range(10):
                    o = ctx.apply_M(F, r)
                    bad += int(not np.array_equal(o, ref))
                if bad: print(precision, n, wall, "variant", v, "nondeterministic runs:", bad, "maxdiff", np.abs(o - ref).max())
            ctx.close()
print("done")
