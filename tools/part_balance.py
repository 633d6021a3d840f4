"""Developer probe (one GPU): kernel time of every share of an n_parts-way partition of the unit triangle,
with equal-count (w = 1) and equal-cost (default w) cuts -- the load imbalance a partitioned suspension
sees, measured without the other GPUs.  One JSON line per (precision, wall, weight)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rigid_body_light_b200._lib import Context  # noqa: E402
from rigid_body_light_b200.shells import sphere_suspension  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bodies", type=int, default=1000)
    ap.add_argument("--shell", type=int, default=162)
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--precisions", default="single,double")
    ap.add_argument("--walls", default="1")
    ap.add_argument("--weights", default="1.0,0,0.7,0.85")
    args = ap.parse_args()
    for precision in args.precisions.split(","):
        tdt = torch.float32 if precision == "single" else torch.float64
        for wall in [bool(int(w)) for w in args.walls.split(",")]:
            ctx = Context(precision)
            s = sphere_suspension(args.bodies, args.shell, wall)
            ctx.set_parameters(s["a"], 0.01, 1.0, 1.0, s["cfg"] - s["cfg"].mean(axis=0))
            ctx.set_flags(0, wall)
            ctx.set_config(s["X"], s["Q"])
            n = args.bodies * args.shell
            r = torch.empty(3 * n, dtype=tdt, device="cuda")
            ctx.call("rbl_dev_blob_positions", r.data_ptr())
            F = torch.randn(3 * n, dtype=tdt, device="cuda")
            out = torch.empty(3 * n, dtype=tdt, device="cuda")
            ctx.call("rbl_sync")
            ctx.call("rbl_profile_matvec", 1)
            for w in [float(x) for x in args.weights.split(",")]:
                ctx.call("rbl_set_split_weight", w)
                times = []
                for parts in (1, args.parts):
                    row = []
                    for p in range(parts):
                        ctx.call("rbl_dev_apply_M_part", F.data_ptr(), r.data_ptr(), n, p, parts, out.data_ptr())
                        ctx.call("rbl_sync")
                        ctx.matvec_profile(reset=True)
                        for _ in range(args.reps):
                            ctx.call("rbl_dev_apply_M_part", F.data_ptr(), r.data_ptr(), n, p, parts, out.data_ptr())
                        ms, _ = ctx.matvec_profile(reset=True)
                        row.append(ms)
                    times.append(row)
                whole, shares = times[0][0], times[1]
                print(json.dumps({"precision": precision, "wall": wall, "n": n, "weight": w if w > 0 else "default",
                                  "whole_ms": round(whole, 4), "share_ms": [round(t, 4) for t in shares],
                                  "max_share_ms": round(max(shares), 4), "ideal_share_ms": round(whole / args.parts, 4),
                                  "efficiency_kernel_only": round(whole / args.parts / max(shares), 4)}), flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
