#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the RPY mobility hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): 1000 spheres
of shell_N_162 above a wall = 162 000 blobs; one STEP = one application of the saddle
operator [M lam - K U ; K^T lam], i.e. one wall-corrected RPY mobility product B M B lam over
N^2 = 2.6244e10 ordered blob pairs plus the K / K^T products.  Metric: ordered blob-pair
interactions per second (whole job).  Strong scaling: bodies are partitioned over the ranks
(rbl_comm_init: the collective saddle operator of the library), lambda is gathered each step, the pair
work is split in equal shares and the partial products are reduce-scattered -- both exchanges by the
library's own kernels over NVLink peer memory (csrc/rbl_peer.cuh), self-checked against the NCCL
collectives before anything is timed and replaced by them where peer memory is unavailable (`exchange`).
`bd_step`: BASELINE.json's second metric, one full fluctuating rigid BD step on configs[2]'s suspension
partitioned over the ranks, seconds per step.

`value`   : device-resident inputs, CUDA events on the launching stream, max over ranks.
`e2e`     : the same step through the host-buffer C ABI (rbl_apply_saddle at N=1; pinned
            host -> device -> sharded step -> host at N>1), copies inside the timed region.
`roofline`: the product kernel alone (events recorded around every launch inside the timed
            region) against the FMA-pipe peak measured live by a microbenchmark.  The bound
            is FP32 (FP64) CUDA-core issue, not HBM and not tensor cores.
`parity`  : after the timed region, at EVERY GPU count, sampled output rows of this run (first / last blob,
            both sides of every rank boundary, seeded random rows; --parity-rows, default 64) against the CPU
            oracle on the inputs the GPU saw; raises above 1e-5 (float) / 1e-12 (double).  The oracle is the
            checker here, outside every timed region.
`bd_step` : one warm-up + three timed full BD steps (min / mean / max), then one untimed profiled step
            (wall clock per phase, summed product-kernel time); the double step also with mixed
            precision modes 1 and 2 (include/rbl.h rbl_set_mixed_precision).
`cpu_baseline` / --impl reference: the reference's own dense 3N x 3N assembly + GEMV
            (rotne_prager_tensor / make_damp_mat / apply_M, c_rigid_obj.cpp:413-459,618-659, compiled
            from the reference source into oracle/_ref/libref_apply_M.so; the oracle's port where that
            library is absent), single thread like the reference, on a bounded sample of the same
            suspension.  The only place bench.py touches oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rpy_blob_pair_interactions_per_s"
UNIT = "pairs/s"
FLOPS_PER_PAIR = {True: 127.0, False: 35.0}  # SURVEY.md section 8d convention
WORKLOADS = {
    # name: (bodies, shell, wall)
    "cfg2": (1000, 162, True),    # BASELINE.json configs[1]  (default, the bench line)
    "cfg3": (4096, 42, True),     # configs[2] geometry
    "cfg4": (1000, 2562, False),  # configs[3]
    "cfg5": (10000, 642, True),   # configs[4]
    "cfg5s": (200, 642, True),    # configs[4]'s body (shell_N_642) at 1/50 of its size
    "small": (64, 42, True),      # CI-sized
}
NVSMI_FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                "clocks_event_reasons.sw_power_cap")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = None
        self.p = None

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={NVSMI_FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if c[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def build_suspension(workload):
    from rigid_body_light_b200.shells import sphere_suspension

    nb, shell, wall = WORKLOADS[workload]
    s = sphere_suspension(nb, shell, wall)
    s["wall"] = wall
    s["n_bodies"], s["n_blb"] = nb, shell
    return s



PARITY_TOL = {"single": 1e-5, "double": 1e-12}  # BASELINE.json north_star


def oracle_parity(precision, op, s, x_local_np, out_local_np, ranges, rank, world, n_rows, dist, torch):
    """Sampled rows of THIS run's output against the CPU oracle (outside every timed region; the
    oracle is the checker here, never the thing measured).  Rows: the first and the last blob of
    the suspension, the two blobs either side of every rank boundary, and seeded random rows, up
    to `n_rows`; every rank checks the sampled rows it owns -- slip rows (M lam - K U)_i against
    oracle.apply_M(rows) - oracle.K_dot, and the force/torque rows K^T lam of the bodies those
    blobs belong to -- on the same inputs the GPU saw (positions as placed on the device, lambda
    and U in the run's precision).  Raises if the relative L2 error over the sampled rows exceeds
    north_star's bound (1e-5 float, 1e-12 double)."""
    from oracle import oracle as orc

    ndt = np.float32 if precision == "single" else np.float64
    nb, n_blb, wall = s["n_bodies"], s["n_blb"], s["wall"]
    n_all = nb * n_blb
    lo, hi = ranges[rank]
    t0, n_local = lo * n_blb, (hi - lo) * n_blb
    rows = {0, n_all - 1}
    for (l, _h) in ranges[1:]:
        rows.update((l * n_blb - 1, l * n_blb))
    extra = np.random.default_rng(7).permutation(n_all)
    for r in extra:
        if len(rows) >= n_rows:
            break
        rows.add(int(r))
    rows = np.array(sorted(rows), dtype=np.int64)
    mine = rows[(rows >= t0) & (rows < t0 + n_local)]
    r_all = op.r_all.cpu().numpy().astype(np.float64)
    lam_all = op.lam_all.cpu().numpy().astype(np.float64)  # the all-gathered lambda of the last step
    X_local = np.asarray(s["X"][lo:hi], dtype=ndt).astype(np.float64)
    U_local = x_local_np[3 * n_local:].astype(ndt).astype(np.float64).reshape(-1, 6)
    got = out_local_np.astype(np.float64)
    num = den = knum = kden = 0.0
    worst = 0.0
    if mine.size:
        want = orc.apply_M(lam_all, r_all, s["a"], 1.0, wall, rows=mine.astype(np.int32)).reshape(-1, 3)
        b_of = (mine - t0) // n_blb
        rho = r_all.reshape(-1, 3)[mine] - X_local[b_of]
        want = want - (U_local[b_of, :3] + np.cross(U_local[b_of, 3:], rho))
        g = got[: 3 * n_local].reshape(-1, 3)[mine - t0]
        num, den = float(((g - want) ** 2).sum()), float((want ** 2).sum())
        worst = float((np.linalg.norm(g - want, axis=1) / np.linalg.norm(want, axis=1)).max())
        bodies = np.unique(b_of)
        rl = r_all.reshape(-1, 3)[t0:t0 + n_local].reshape(-1, n_blb, 3)[bodies]
        ll = lam_all.reshape(-1, 3)[t0:t0 + n_local].reshape(-1, n_blb, 3)[bodies]
        wantk = orc.KT_dot(ll.reshape(-1), rl.reshape(-1), X_local[bodies], n_blb).reshape(-1, 6)
        gk = got[3 * n_local:].reshape(-1, 6)[bodies]
        knum, kden = float(((gk - wantk) ** 2).sum()), float((wantk ** 2).sum())
    acc = [num, den, knum, kden, float(mine.size)]
    if world > 1:
        t = torch.tensor(acc, dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        w = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        acc, worst = t.tolist(), float(w.item())
    rel = float(np.sqrt(acc[0] / acc[1]))
    rel_k = float(np.sqrt(acc[2] / acc[3]))
    tol = PARITY_TOL[precision]
    if not (rel <= tol and rel_k <= tol):
        raise AssertionError(f"oracle parity FAILED ({precision}, {world} GPUs): slip rows {rel:.3e}, K^T rows {rel_k:.3e} > {tol:g}")
    return {"rows": int(acc[4]), "rel_err": rel, "max_row_rel_err": worst, "kt_rel_err": rel_k, "tol": tol,
            "rows_include": "first/last blob, both sides of every rank boundary, seeded random rows",
            "against": "oracle.apply_M(rows=...) - oracle.K_dot and oracle.KT_dot (CPU restatement of c_rigid_obj.cpp:31-142,"
                       "404-459,618-659; long-double row sums) on the positions the device placed; relative L2 over the sampled rows"}

# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def _protect_stdout():
    """The contract is ONE JSON line on stdout, but native libraries write there too (NCCL prints
    its version banner on the first communicator).  Point fd 1 at stderr for the run and return a
    file on the ORIGINAL stdout for the result line."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def run_ours(args):
    result_out = _protect_stdout()
    import torch
    import torch.distributed as dist

    from rigid_body_light_b200._lib import Context
    from rigid_body_light_b200.sharding import CudaShard, PartitionedRigidBody, ShardedSaddle, body_ranges, slice_system

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    s = build_suspension(args.workload)
    nb, n_blb, wall = s["n_bodies"], s["n_blb"], s["wall"]
    n_all = nb * n_blb
    ranges = body_ranges(nb, world)
    lo, hi = ranges[rank]
    ref = s["cfg"] - s["cfg"].mean(axis=0)
    vec = np.random.default_rng(2).standard_normal(3 * n_all + 6 * nb)
    x_local_np = slice_system(vec, ranges, n_blb, rank)
    pairs = float(n_all) * float(n_all)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    results = {}
    for precision in (["single", "double"] if args.dtype == "both" else [args.dtype]):
        tdt = torch.float32 if precision == "single" else torch.float64
        ndt = np.float32 if precision == "single" else np.float64
        pb = None
        exchange = None
        if world == 1:
            ctx = Context(precision, device=local_rank)
            ctx.set_parameters(s["a"], 0.01, 1.0, 1.0, ref)
            ctx.set_flags(0, int(wall))
            ctx.set_config(s["X"][lo:hi], s["Q"][lo:hi])
        else:
            # the collective saddle operator of the library (rbl_comm_init): the exchanges around the product are
            # its own peer-memory kernels where every rank could map every other rank's buffer, NCCL otherwise
            pb = PartitionedRigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=wall, block_PC=False,
                                      precision=precision, rank=rank, world=world, dist=dist, device=local_rank)
            ctx = pb.ctx
        shard = CudaShard(ctx, (hi - lo) * n_blb, tdt)  # (puts the context on torch's current stream)
        op = ShardedSaddle(shard, nb, n_blb, rank, world, dist if world > 1 else None)
        x_local = torch.from_numpy(x_local_np.astype(ndt)).cuda()
        out_local = torch.empty_like(x_local)
        op.refresh_positions()
        peak = max(ctx.fma_peak(20000) for _ in range(3))  # TFLOP/s, live, this GPU, this precision

        if world == 1:
            apply_op = lambda: op.apply(x_local, out_local)  # noqa: E731
        else:
            apply_op = lambda: pb.apply_saddle_dev(x_local.data_ptr(), out_local.data_ptr())  # noqa: E731
            exchange = {"mode": pb.exchange, "why_not_peer": pb.exchange_why or None}
            if pb.exchange == "peer":
                # self-check before anything is timed: the same operator through both protocols; a failure of the
                # peer exchange (error code, or a result that differs beyond rounding) leaves NCCL in charge
                ok = 1.0
                try:
                    apply_op()
                    ctx.call("rbl_sync")
                    a_peer = out_local.clone()
                    pb.set_exchange("nccl")
                    apply_op()
                    ctx.call("rbl_sync")
                    den = float(torch.linalg.vector_norm(out_local.double()))
                    dev = float(torch.linalg.vector_norm(a_peer.double() - out_local.double())) / max(den, 1e-300)
                    exchange["selfcheck_rel_diff_vs_nccl"] = dev
                    if not dev < (2e-6 if precision == "single" else 1e-13):
                        ok = 0.0
                except Exception as exc:  # noqa: BLE001
                    exchange["selfcheck_error"] = str(exc)[:200]
                    ok = 0.0
                t = torch.tensor([ok], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                if float(t.item()) > 0.5:
                    pb.set_exchange("peer")
                else:
                    try:
                        pb.set_exchange("nccl")
                    except Exception:  # noqa: BLE001
                        pass
                exchange["mode"] = pb.exchange

        def step():
            ctx.call("rbl_flush_l2")  # inputs (5 MB) are smaller than the 126 MB L2: flush between steps
            apply_op()

        for _ in range(args.warmup):
            step()
        barrier()
        ctx.call("rbl_profile_matvec", 1)
        ctx.matvec_profile(reset=True)
        launches0 = ctx.launch_count()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            ctx.call("rbl_comm_profile", None, None, 1)  # clear
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        comm = None
        if world > 1:
            import ctypes

            ms3, cn = (ctypes.c_double * 3)(), ctypes.c_int()
            ctx.call("rbl_comm_profile", ms3, ctypes.byref(cn), 1)
            comm = [max_over_ranks(float(v)) for v in ms3]
        clocks = sampler.stop() if rank == 0 else None
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = ctx.launch_count() - launches0
        kern_ms, kern_n = ctx.matvec_profile(reset=True)
        ctx.call("rbl_profile_matvec", 0)
        kern_ms = max_over_ranks(kern_ms)
        ms_step = ms_total / args.steps
        if world > 1 and exchange["mode"] == "peer":
            # A/B in the same process, AFTER the loop that `value` reports: the same timed loop with the NCCL
            # collectives in place of the peer-memory exchange, then with the peer-memory exchange once more (the
            # repeat shows how much of any difference to `value` is the order of the loops, not the protocol;
            # tools/exchange_ab.py alternates the two on the same contexts)
            def timed_loop(mode):
                pb.set_exchange(mode)
                for _ in range(args.warmup):
                    step()
                barrier()
                ctx.call("rbl_profile_matvec", 1)  # the same event records between the kernels as in the loop above
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(args.steps):
                    step()
                f1.record()
                barrier()
                ctx.matvec_profile(reset=True)
                ctx.call("rbl_comm_profile", None, None, 1)
                ctx.call("rbl_profile_matvec", 0)
                return max_over_ranks(f0.elapsed_time(f1)) / args.steps

            exchange["ms_per_step_with_nccl_collectives"] = timed_loop("nccl")
            exchange["ms_per_step_peer_repeated_after_nccl"] = timed_loop("peer")

        # end to end: host buffers, copies inside the timed region
        nbytes = x_local.numel() * x_local.element_size()
        if world == 1:
            import ctypes

            hx, ho = ctypes.c_void_p(), ctypes.c_void_p()
            ctx.call("rbl_pinned_alloc", nbytes, ctypes.byref(hx))
            ctx.call("rbl_pinned_alloc", nbytes, ctypes.byref(ho))
            x_host = np.ascontiguousarray(x_local_np.astype(ndt))  # keep alive across the memmove
            ctypes.memmove(hx, x_host.ctypes.data, nbytes)
            for _ in range(args.e2e_warmup):
                ctx.call("rbl_apply_saddle", hx, ho)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ctx.call("rbl_flush_l2")              # same experiment as `value`: L2 flushed between steps
                ctx.call("rbl_apply_saddle", hx, ho)  # H2D + step + D2H, synchronous
            e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            host_out = np.frombuffer((ctypes.c_char * nbytes).from_address(ho.value), dtype=ndt).copy()
            ctx.call("rbl_pinned_free", hx)
            ctx.call("rbl_pinned_free", ho)
        else:
            hx = torch.from_numpy(x_local_np.astype(ndt)).pin_memory()
            ho = torch.empty_like(hx).pin_memory()
            for _ in range(args.e2e_warmup):
                x_local.copy_(hx, non_blocking=True); apply_op(); ho.copy_(out_local, non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ctx.call("rbl_flush_l2")
                x_local.copy_(hx, non_blocking=True)
                apply_op()
                ho.copy_(out_local, non_blocking=True)
                torch.cuda.synchronize()
            e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
            host_out = ho.numpy().copy()
        dev_out = out_local.cpu().numpy()
        # the symmetric kernel accumulates with floating-point atomics: the two paths agree to
        # rounding, not bit for bit
        dd = np.abs(dev_out.astype(np.float64) - host_out.astype(np.float64))
        agree = float(np.linalg.norm(dd) / np.linalg.norm(host_out.astype(np.float64)))
        if not agree < (2e-6 if precision == "single" else 1e-13):
            raise AssertionError(f"host-buffer and device-resident paths disagree ({precision}): relative L2 {agree:.3e}, "
                                 f"max |diff| {dd.max():.3e} at {int(dd.argmax())}")
        ctx.call("rbl_sync")
        parity = None
        if world > 1:  # the all-gathered lambda the parity check reads (the library keeps its own copy inside)
            op._allgather(x_local[: 3 * op.n_local], op.lam_all)
        if args.parity_rows > 0:
            parity = oracle_parity(precision, op, s, x_local_np, dev_out, ranges, rank, world, args.parity_rows,
                                   dist if world > 1 else None, torch)

        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get(f"{args.workload}/{precision}/{world}")
        except Exception:
            pass
        alg_tflops = (float((hi - lo) * n_blb) * n_all) * FLOPS_PER_PAIR[wall] / (kern_ms * 1e-3) / 1e12
        results[precision] = {
            "value": pairs / (ms_step * 1e-3), "ms_per_step": ms_step, "gpu_launches": int(launches),
            "e2e": {"value": pairs / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(nbytes) * world, "d2h_bytes_per_step": int(nbytes) * world},
            "roofline": {"bound": "fp32_cuda_core" if precision == "single" else "fp64_cuda_core",
                         "kernel": "rbl::rpy_matvec_sym_kernel", "achieved": alg_tflops, "peak": peak,
                         "unit": "TFLOP/s", "frac": alg_tflops / peak, "traffic": traffic,
                         "traffic_note": "DRAM bytes per launch from the committed ncu --set full capture (profiles/); "
                                         "algorithmic bytes = packed records, 162000 x 32 B (fp32) / 64 B (fp64)",
                         "kernel_ms": kern_ms, "kernel_launches_timed": int(kern_n),
                         "algorithmic_flops_per_pair": FLOPS_PER_PAIR[wall],
                         "convention": "SURVEY.md 8d: N^2 ordered pairs x 127 flop (35 free space) / kernel time. The "
                                       "kernel evaluates each UNORDERED pair once (like the reference's i<=j loop) and "
                                       "applies the block in both directions, so it executes ~28% fewer instructions "
                                       "than the convention assumes; issue-slot utilisation is in profiles/",
                         "peak_source": "FMA-chain microbenchmark run live on this GPU (rbl_fma_peak); nominal "
                                        + ("74.4" if precision == "single" else "37.2") + " TFLOP/s at 148 SM x 1.965 GHz"},
            "clocks": clocks, "checksum": float(np.abs(dev_out.astype(np.float64)).sum()), "parity": parity,
            "comm_ms_per_step": None if comm is None else {"allgather_lambda": comm[0], "product_incl_pack": comm[1],
                                                            "reduce_partials": comm[2],
                                                            "note": "CUDA events per step, max over ranks; a rank that "
                                                                    "finishes its share early waits inside the exchange"},
            "exchange": exchange,
        }
        if pb is not None:
            barrier()
            pb.close()  # collective: the peer buffers are unmapped in step
        else:
            ctx.close()
        del op, shard, x_local, out_local

    if world > 1 and any("selfcheck_error" in (r["exchange"] or {}) or
                         ((r["exchange"] or {}).get("selfcheck_rel_diff_vs_nccl") is not None and (r["exchange"] or {}).get("mode") != "peer")
                         for r in results.values()):
        # the peer-memory exchange failed its self-check on this box (every rank saw the same verdict): the
        # BD leg's contexts use the NCCL collectives too
        os.environ["RBL_PEER_EXCHANGE"] = "0"
    bd = None
    if args.bd_steps > 0:
        bd = bd_step_leg(args, rank, world, dist if world > 1 else None, local_rank,
                         ["single", "double"] if args.dtype == "both" else [args.dtype])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_sample(args.workload, "single" if args.dtype != "double" else "double", budget_s=12.0)

    if rank == 0:
        head_p = "single" if args.dtype in ("both", "single") else "double"
        head = results[head_p]
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if head_p == "single" else "f64",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {nb} spheres of shell_N_{n_blb} "
                                   f"{'above a wall' if wall else 'in free space'} = {n_all} blobs; step = apply_saddle "
                                   f"(wall-corrected RPY matvec + K + K^T)",
                       "pairs_per_step": pairs, "parallelism": f"x{world}: bodies in contiguous ranges; lambda gathered and the partial products reduce-scattered "
                                      f"by the library's own kernels over NVLink peer memory (NCCL collectives where "
                                      f"that is unavailable: see `exchange`), equal shares of the unordered-pair tile triangle per rank",
                       "l2": "256 MiB memset between steps inside the timed region (inputs < L2)",
                       "seeds": {"geometry": 0, "quaternions": 1, "vectors": 2}},
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "clocks": head["clocks"], "cpu_baseline": cpu, "comm_ms_per_step": head["comm_ms_per_step"],
            "exchange": head["exchange"], "parity": head["parity"], "bd_step": bd,
        }
        if "double" in results and head_p == "single":
            d = results["double"]
            line["f64"] = {k: d[k] for k in ("value", "ms_per_step", "e2e", "roofline", "gpu_launches", "clocks", "comm_ms_per_step", "exchange", "parity")}
        print(json.dumps(line), file=result_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



def bd_step_leg(args, rank, world, dist, local_rank, precisions):
    """BASELINE.json's second metric: wall time of one full fluctuating rigid BD step (2 Lanczos
    M^{1/2}W, RFD drift, midpoint K/PC, preconditioned GMRES, evolve) through the host-buffer C ABI
    (rbl_bd_step: host arrays in, rigid velocities out), on configs[2]'s suspension partitioned
    over the ranks.  One untimed warm-up step (allocations, module load), then --bd-steps timed
    steps, max over ranks; every step draws fresh noise and moves the bodies."""
    import ctypes

    import torch

    from rigid_body_light_b200.sharding import PartitionedRigidBody

    s = build_suspension(args.bd_workload)
    nb, n_blb, wall = s["n_bodies"], s["n_blb"], s["wall"]
    n3 = 3 * nb * n_blb
    F_ext = np.tile(np.array([0, 0, -1.0, 0, 0, 0]), nb)
    out = {"workload": f"{args.bd_workload}: {nb} spheres of shell_N_{n_blb} {'above a wall' if wall else 'in free space'} "
                       f"= {nb * n_blb} blobs; kBT = 0.0041, dt = 0.01, gravity on every body, block-diagonal PC, block-Cholesky preconditioned paired Lanczos noise",
           "unit": "s/step", "higher_is_better": False, "n_gpus": world, "steps": args.bd_steps,
           "warmup_steps": 1 if args.bd_warmup else 0, "exchange": None}
    legs = [(p, 0) for p in precisions]
    if "double" in precisions and args.bd_mixed:
        legs += [("double", 1), ("double", 2)]  # mixed precision (the float mirror shares the communicator at N > 1)
    for precision, mixed in legs:
        tol, ltol = (1e-4, 1e-4) if precision == "single" else (1e-8, 1e-6)
        pb = PartitionedRigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=wall, block_PC=True,
                                  precision=precision, rank=rank, world=world, dist=dist, device=local_rank)
        if mixed:
            pb.set_mixed_precision(mixed)
        if world > 1:
            out["exchange"] = pb.exchange
        rng = np.random.default_rng(3)
        times, iters, lz, rel = [], [], [], []
        prod0 = 0
        warm = 1 if args.bd_warmup else 0
        for k in range(warm + args.bd_steps):
            noise = tuple(pb.slice_blobs(rng.standard_normal(n3)) for _ in range(3))
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            if k == warm:
                prod0 = int(pb.ctx.L.rbl_product_count(pb.ctx.h))
            t0 = time.perf_counter()
            U, it, rr = pb.bd_step(pb.slice_bodies(F_ext), kBT=0.0041, noise_local=noise, tol=tol, restart=60,
                                   max_iter=args.bd_gmres_max_iter, lanczos_tol=ltol, lanczos_max_iter=args.bd_lanczos_max_iter)
            dt = time.perf_counter() - t0
            if k < warm:
                continue
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            l1, l2 = ctypes.c_int(), ctypes.c_int()
            pb.ctx.L.rbl_bd_stats(pb.ctx.h, ctypes.byref(l1), ctypes.byref(l2))
            times.append(dt); iters.append(int(it)); lz.append([l1.value, l2.value]); rel.append(float(rr))
        products = (int(pb.ctx.L.rbl_product_count(pb.ctx.h)) - prod0) / max(1, args.bd_steps)
        profile_step = None
        if args.bd_profile_step:
            # one more step, NOT timed, with the profiling hooks on: where a step's wall clock goes
            # (a stream sync ends every phase; CUDA events bracket every product kernel)
            pb.ctx.call("rbl_profile_matvec", 1)
            pb.ctx.matvec_profile(reset=True)
            ph = (ctypes.c_double * 6)()
            pb.ctx.call("rbl_bd_phase_ms", ph, 1)
            noise = tuple(pb.slice_blobs(rng.standard_normal(n3)) for _ in range(3))
            t0 = time.perf_counter()
            pb.bd_step(pb.slice_bodies(F_ext), kBT=0.0041, noise_local=noise, tol=tol, restart=60,
                       max_iter=args.bd_gmres_max_iter, lanczos_tol=ltol, lanczos_max_iter=args.bd_lanczos_max_iter)
            prof_s = time.perf_counter() - t0
            kms, kn = pb.ctx.matvec_profile(reset=True)
            pb.ctx.call("rbl_bd_phase_ms", ph, 1)
            pb.ctx.call("rbl_profile_matvec", 0)
            profile_step = {"seconds": prof_s, "product_kernel_seconds": kms * kn * 1e-3, "product_kernel_launches": int(kn),
                            "phase_seconds": dict(zip(["inputs_noise", "lanczos", "rfd", "midpoint", "gmres_incl_pc_build", "evolve_output"],
                                                      [float(v) * 1e-3 for v in ph])),
                            "note": "rank 0, untimed extra step with a stream sync after every phase"}
        X, _ = pb.get_config()
        out[precision if not mixed else f"double_mixed{mixed}"] = {"seconds_per_step": float(np.mean(times)), "seconds_per_step_min": float(np.min(times)),
                          "seconds_per_step_max": float(np.max(times)), "seconds_each_step": [float(t) for t in times],
                          "gmres_iterations": iters, "lanczos_iterations": lz,
                          "gmres_tol": tol, "lanczos_tol": ltol, "gmres_max_iter": args.bd_gmres_max_iter,
                          "lanczos_max_iter": args.bd_lanczos_max_iter, "relres": rel, "mobility_products_per_step": products,
                          "min_body_height_after": float(X[:, 2].min()), "U_norm_local": float(np.linalg.norm(U)),
                          "profile_step": profile_step,
                          "mixed_precision": {0: None, 1: "float GMRES corrections, double-residual refinement (same relres <= tol)",
                                              2: "mode 1 + float mobility products inside the Lanczos noise"}[mixed]}
        pb.close()
    return out

# --------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port), bounded sample
# --------------------------------------------------------------------------------------
def _sample_inputs(workload, n_bodies_sample, ndt):
    from oracle import oracle as orc

    s = build_suspension(workload)
    nb = min(n_bodies_sample, s["n_bodies"])
    ref = orc.remove_mean(s["cfg"])
    r = orc.blob_positions(s["X"][:nb], s["Q"][:nb], ref).astype(ndt)
    F = np.random.default_rng(2).standard_normal(r.size).astype(ndt)
    return s, nb, r, F


def cpu_reference_sample(workload, precision, budget_s, steps=1, warmup=0):
    """Times apply_M exactly as the reference runs it (dense assembly + GEMV, 1 thread) on the
    first bodies of the workload, sized so one step takes about budget_s/(steps+warmup)."""
    from oracle import oracle as orc

    ndt = np.float32 if precision == "single" else np.float64
    nb_total, n_blb, wall = WORKLOADS[workload]
    # the reference's OWN members (rotne_prager_tensor + make_damp_mat + apply_M compiled from its
    # source into oracle/_ref/libref_apply_M.so) when that library is there, else the oracle's port
    use_ref = orc.ref_apply_M_lib() is not None

    def dense_apply(F_, r_, a_):
        if use_ref:
            return orc.ref_apply_M(F_, r_, a_, 1.0, wall, dtype=ndt)
        return orc.apply_M_dense(F_, r_, a_, 1.0, wall, dtype=ndt)

    # calibrate on a small piece
    s, nb, r, F = _sample_inputs(workload, max(1, 600 // n_blb), ndt)
    t0 = time.perf_counter()
    dense_apply(F, r, s["a"])
    rate = (r.shape[0] ** 2) / max(time.perf_counter() - t0, 1e-6)
    per_step = budget_s / max(1, steps + warmup)
    n_target = int(np.sqrt(rate * per_step))
    mem_cap = int(np.sqrt(6e9 / (9 * np.dtype(ndt).itemsize)))  # dense matrix <= 6 GB
    nb_s = max(1, min(nb_total, min(n_target, mem_cap) // n_blb))
    s, nb, r, F = _sample_inputs(workload, nb_s, ndt)
    n = r.shape[0]
    for _ in range(warmup):
        dense_apply(F, r, s["a"])
    t0 = time.perf_counter()
    for _ in range(steps):
        dense_apply(F, r, s["a"])
    dt = (time.perf_counter() - t0) / steps
    # best-effort CPU: matrix-free OpenMP oracle on all host cores, sampled target rows of the
    # FULL workload (float64 with long-double row sums; a courtesy number, not the reference)
    sf, nbf, rf, Ff = _sample_inputs(workload, nb_total, np.float64)
    rows = np.random.default_rng(5).choice(rf.shape[0], min(rf.shape[0], 512), replace=False)
    t0 = time.perf_counter()
    orc.apply_M(Ff, rf, sf["a"], 1.0, wall, rows=rows)
    dt_mf = time.perf_counter() - t0
    return {
        "value": n * n / dt, "unit": UNIT, "cores": 1, "kind": "reference" if use_ref else "port",
        "sample": f"first {nb} bodies of {workload} ({n} blobs, {precision}): dense {3*n}x{3*n} assembly + GEMV "
                  + ("by the reference's own rotne_prager_tensor / make_damp_mat / apply_M members compiled from its source "
                     "(c_rigid_obj.cpp:413-459,618-659; Eigen's dense ops supplied by oracle/eigen_shim.inc), "
                     if use_ref else "like c_rigid_obj.cpp:413-459,641-659 (oracle port), ")
                  + f"single thread (the reference has no threading); "
                  f"the full workload would need {9 * (nb_total * n_blb) ** 2 * np.dtype(ndt).itemsize / 1e9:.0f} GB",
        "ms_per_step": dt * 1e3,
        "best_effort_all_cores": {"value": rows.size * rf.shape[0] / dt_mf, "unit": UNIT, "cores": orc.num_threads(),
                                  "kind": "port", "sample": f"matrix-free OpenMP oracle, {rows.size} sampled target rows "
                                                            f"x {rf.shape[0]} sources of the full workload, float64"},
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    precision = "single" if args.dtype in ("both", "single") else "double"
    nb, n_blb, wall = WORKLOADS[args.workload]
    budget = float(os.environ.get("RBL_BENCH_CPU_BUDGET_S", "100"))  # whole-run CPU budget (tests shrink it)
    cpu = cpu_reference_sample(args.workload, precision, budget_s=budget, steps=args.steps, warmup=args.warmup)
    n_all = nb * n_blb
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32" if precision == "single" else "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {nb} spheres of shell_N_{n_blb} "
                               f"{'above a wall' if wall else 'in free space'} = {n_all} blobs; step = apply_M on a bounded "
                               f"sample (the reference's dense algorithm cannot hold the full workload)",
                   "sample": cpu["sample"]},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="both", choices=["both", "single", "double"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-rows", type=int, default=64,
                    help="sampled output rows checked against the CPU oracle after the timed region, at every N (0 = off)")
    ap.add_argument("--bd-steps", type=int, default=3, help="timed full BD steps after the matvec bench (0 = skip)")
    ap.add_argument("--bd-workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--bd-warmup", type=int, default=1, help="0: time the very first BD step (allocations included)")
    ap.add_argument("--e2e-warmup", type=int, default=-1, help="warm-up steps of the host-buffer leg (default: --warmup; the "
                    "kernels are warm from the device-resident leg, so 1 is enough for the seconds-long cfg5 products)")
    ap.add_argument("--bd-profile-step", type=int, default=1, help="0: skip the extra, untimed, profiled BD step")
    ap.add_argument("--bd-mixed", type=int, default=1, help="also time the double BD step with mixed precision modes 1 and 2 (N=1)")
    ap.add_argument("--bd-gmres-max-iter", type=int, default=200)
    ap.add_argument("--bd-lanczos-max-iter", type=int, default=80)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.e2e_warmup = max(1, args.warmup if args.e2e_warmup < 0 else args.e2e_warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
