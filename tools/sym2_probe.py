"""One two-right-hand-side product at BASELINE configs[2] size with the DEFAULT variant (ncu target)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from Rigid import RigidBody  # noqa: E402
from rigid_body_light_b200.shells import sphere_suspension  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "single"
s = sphere_suspension(4096, 42, True)
cb = RigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, precision=precision)
r = cb.get_blob_positions()
rng = np.random.default_rng(0)
F1, F2 = rng.standard_normal(r.size), rng.standard_normal(r.size)
for _ in range(2):
    o1, o2 = cb.apply_M2(F1, F2, r)
print("ok", float(np.abs(o1).sum()))
