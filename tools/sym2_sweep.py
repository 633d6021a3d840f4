"""Sweeps the two-right-hand-side kernel variants against two single-vector products and times
the BD step with and without the paired Lanczos (GPU box).  One JSON line per measurement."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_light_b200._lib import Context  # noqa: E402
from rigid_body_light_b200.shells import sphere_suspension  # noqa: E402


def main():
    import torch

    which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    nb, shell, wall = {"cfg2": (1000, 162, True), "cfg3": (4096, 42, True), "cfg4s": (200, 2562, False)}[which]
    s = sphere_suspension(nb, shell, wall)
    ref = s["cfg"] - s["cfg"].mean(axis=0)
    n = nb * shell
    for precision in ("single", "double"):
        tdt = torch.float32 if precision == "single" else torch.float64
        ctx = Context(precision)
        ctx.set_parameters(s["a"], 0.01, 1.0, 1.0, ref)
        ctx.set_flags(0, int(wall))
        ctx.set_config(s["X"], s["Q"])
        r = torch.empty(3 * n, dtype=tdt, device="cuda")
        ctx.call("rbl_dev_blob_positions", r.data_ptr())
        g = torch.Generator(device="cuda").manual_seed(1)
        F1 = torch.randn(3 * n, dtype=tdt, device="cuda", generator=g)
        F2 = torch.randn(3 * n, dtype=tdt, device="cuda", generator=g)
        o1, o2, s1, s2 = (torch.empty_like(F1) for _ in range(4))
        ctx.call("rbl_sync")

        def timed(fn, reps=3):
            fn(); ctx.call("rbl_sync")
            ctx.timer_start()
            for _ in range(reps):
                fn()
            return ctx.timer_stop() / reps

        t_single = timed(lambda: (ctx.call("rbl_dev_apply_M", F1.data_ptr(), r.data_ptr(), n, 0, n, s1.data_ptr()),
                                  ctx.call("rbl_dev_apply_M", F2.data_ptr(), r.data_ptr(), n, 0, n, s2.data_ptr())))
        nv = ctx.L.rbl_num_sym2_variants(ctx.h)
        for v in range(nv):
            T, NT = ctypes.c_int(), ctypes.c_int()
            ctx.L.rbl_sym2_variant_info(ctx.h, v, ctypes.byref(T), ctypes.byref(NT))
            ctx.call("rbl_set_sym2_variant", v)
            t2 = timed(lambda: ctx.call("rbl_dev_apply_M2", F1.data_ptr(), F2.data_ptr(), r.data_ptr(), n, o1.data_ptr(), o2.data_ptr()))
            err = max(float((o1 - s1).norm() / s1.norm()), float((o2 - s2).norm() / s2.norm()))
            print(json.dumps({"workload": which, "precision": precision, "variant": [T.value, NT.value], "two_rhs_ms": t2,
                              "two_single_ms": t_single, "speedup": t_single / t2, "rel_diff": err}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
