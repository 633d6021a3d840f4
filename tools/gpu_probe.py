"""Developer probe (GPU box): FMA peaks, every matvec variant at the benchmark size.
Not part of the product or the bench contract; prints one JSON line per measurement."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rigid_body_light_b200._lib import Context  # noqa: E402
from rigid_body_light_b200.shells import sphere_suspension  # noqa: E402

FLOPS = {True: 127.0, False: 35.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bodies", type=int, default=1000)
    ap.add_argument("--shell", type=int, default=162)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--precisions", default="single,double")
    ap.add_argument("--walls", default="1,0")
    ap.add_argument("--variants", default="all")
    ap.add_argument("--kernel", default="symmetric", choices=["symmetric", "ordered"])
    args = ap.parse_args()
    for precision in args.precisions.split(","):
        ctx = Context(precision)
        peak = max(ctx.fma_peak(20000) for _ in range(3))
        print(json.dumps({"probe": "fma_peak", "precision": precision, "tflops": round(peak, 2)}), flush=True)
        tdt = torch.float32 if precision == "single" else torch.float64
        for wall in [bool(int(w)) for w in args.walls.split(",")]:
            s = sphere_suspension(args.bodies, args.shell, wall)
            ref = s["cfg"] - s["cfg"].mean(axis=0)
            ctx.set_parameters(s["a"], 0.01, 1.0, 1.0, ref)
            ctx.set_flags(0, wall)
            ctx.set_config(s["X"], s["Q"])
            n = args.bodies * args.shell
            r = torch.empty(3 * n, dtype=tdt, device="cuda")
            ctx.call("rbl_dev_blob_positions", r.data_ptr())
            F = torch.randn(3 * n, dtype=tdt, device="cuda")
            out = torch.empty(3 * n, dtype=tdt, device="cuda")
            ctx.call("rbl_sync")
            ctx.call("rbl_set_matvec_mode", 1 if args.kernel == "ordered" else 0)
            nv = ctx.L.rbl_num_matvec_variants(ctx.h) if args.kernel == "ordered" else ctx.L.rbl_num_sym_variants(ctx.h)
            vs = range(nv) if args.variants == "all" else [int(v) for v in args.variants.split(",")]
            ctx.call("rbl_profile_matvec", 1)
            for v in vs:
                T, th = ctypes.c_int(), ctypes.c_int()
                if args.kernel == "ordered":
                    ctx.L.rbl_matvec_variant_info(ctx.h, v, ctypes.byref(T), ctypes.byref(th))
                    ctx.call("rbl_set_matvec_variant", v)
                else:
                    ctx.L.rbl_sym_variant_info(ctx.h, v, ctypes.byref(T), ctypes.byref(th))
                    ctx.call("rbl_set_sym_variant", v)
                ctx.call("rbl_dev_apply_M", F.data_ptr(), r.data_ptr(), n, 0, n, out.data_ptr())  # warm-up
                ctx.call("rbl_sync")
                ctx.matvec_profile(reset=True)
                ctx.timer_start()
                for _ in range(args.reps):
                    ctx.call("rbl_dev_apply_M", F.data_ptr(), r.data_ptr(), n, 0, n, out.data_ptr())
                ms = ctx.timer_stop() / args.reps
                kms, nl = ctx.matvec_profile(reset=True)
                rc = ctx.L.rbl_sym_variant_chunk(ctx.h, v) if args.kernel == "symmetric" else 0
                pairs = float(n) * n
                tf = pairs * FLOPS[wall] / (kms * 1e-3) / 1e12
                print(json.dumps({"probe": "matvec", "kernel": args.kernel, "precision": precision, "wall": wall, "variant": v, "T": T.value, "rc": rc,
                                  "threads": th.value, "n": n, "ms_call": round(ms, 3), "ms_kernel": round(kms, 3),
                                  "gpairs_s": round(pairs / (ms * 1e-3) / 1e9, 1), "alg_tflops": round(tf, 2),
                                  "frac_of_fma_peak": round(tf / peak, 3), "sum": float(out.double().abs().sum())}),
                      flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
