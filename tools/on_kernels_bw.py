"""HBM bandwidth of the O(N) rigid-body kernels (placement, K, K^T, diagonal-PC apply) at
BASELINE.json configs[4] size on ONE GPU (10 000 bodies x 642 blobs = 6.42 M blobs: every vector
is 77 MB in fp32, so reads + writes exceed the 126 MB L2 and each launch streams from HBM).
Algorithmic bytes are DESIGN.md section 5's; peak = MEASURED_PEAKS.json hbm_gbs.  One JSON line
per kernel and precision."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rigid_body_light_b200._lib import Context  # noqa: E402
from rigid_body_light_b200.shells import sphere_suspension  # noqa: E402


def main():
    import torch

    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "MEASURED_PEAKS.json"
    except Exception:
        peak, src = 6533.5, "fallback"
    nb, shell = 10000, 642
    s = sphere_suspension(nb, shell, True)
    ref = s["cfg"] - s["cfg"].mean(axis=0)
    n = nb * shell
    for precision in ("single", "double"):
        tdt = torch.float32 if precision == "single" else torch.float64
        sz = 4 if precision == "single" else 8
        ctx = Context(precision)
        ctx.set_parameters(s["a"], 0.01, 1.0, 1.0, ref)
        ctx.set_flags(0, 1)
        ctx.set_config(s["X"], s["Q"])
        g = torch.Generator(device="cuda").manual_seed(1)
        lam = torch.randn(3 * n, dtype=tdt, device="cuda", generator=g)
        U = torch.randn(6 * nb, dtype=tdt, device="cuda", generator=g)
        x = torch.randn(3 * n + 6 * nb, dtype=tdt, device="cuda", generator=g)
        out3 = torch.empty(3 * n, dtype=tdt, device="cuda")
        out6 = torch.empty(6 * nb, dtype=tdt, device="cuda")
        outx = torch.empty_like(x)
        ctx.call("rbl_dev_apply_PC", x.data_ptr(), outx.data_ptr())  # builds the diagonal PC
        ctx.call("rbl_sync")
        cases = {
            # name: (callable, algorithmic bytes)
            "place_blobs": (lambda: ctx.call("rbl_dev_blob_positions", out3.data_ptr()), (3 * n + 7 * nb + 3 * shell) * sz),
            "k_dot": (lambda: ctx.call("rbl_dev_K_dot", U.data_ptr(), out3.data_ptr()), (3 * n + 3 * n + 6 * nb + 3 * nb) * sz),
            "kt_dot": (lambda: ctx.call("rbl_dev_KT_dot", lam.data_ptr(), out6.data_ptr()), (3 * n + 3 * n + 6 * nb + 3 * nb) * sz),
            # y = Minv slip (read 3N + 3N diag, write 3N); finish: read y, Y (6*3N), r, write Lambda
            "apply_PC_diag (2 launches)": (lambda: ctx.call("rbl_dev_apply_PC", x.data_ptr(), outx.data_ptr()),
                                           (3 * 3 * n + (1 + 6 + 1 + 1) * 3 * n) * sz),
        }
        # references on the SAME buffers: a write-only fill and a read+write copy (what the denominators mean)
        for name, op, nbytes in (("reference: fill (write 3N)", lambda: out3.zero_(), 3 * n * sz),
                                 ("reference: copy (read 3N, write 3N)", lambda: out3.copy_(lam), 2 * 3 * n * sz)):
            for _ in range(3):
                op()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                op()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(json.dumps({"kernel": name, "precision": precision, "blobs": n, "ms": ms, "algorithmic_bytes": nbytes,
                              "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "peak_gbs": peak, "frac": nbytes / (ms * 1e-3) / 1e9 / peak}),
                  flush=True)
        for name, (fn, nbytes) in cases.items():
            for _ in range(3):
                fn()
            ctx.call("rbl_sync")
            reps = 20
            ctx.timer_start()
            for _ in range(reps):
                fn()
            ms = ctx.timer_stop() / reps
            gbs = nbytes / (ms * 1e-3) / 1e9
            print(json.dumps({"kernel": name, "precision": precision, "blobs": n, "bodies": nb, "ms": ms,
                              "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "peak_gbs": peak, "peak_source": src,
                              "frac": gbs / peak}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
