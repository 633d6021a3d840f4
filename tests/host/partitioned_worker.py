"""torchrun worker for tests/test_gpu_partitioned.py: every rank holds its body range of a small
suspension (NCCL inside librbl, rbl_comm_init) and checks the collective operators against the
same suspension on ONE context built on the rank's own GPU.  Prints PARTITIONED-OK on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def chk(err, tol, *what):
    """assert err < tol and print the observed value (the tolerances here are held within ~10x of it)"""
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"OBS world={os.environ.get('WORLD_SIZE')} {' '.join(str(w) for w in what)}: {err:.3e} (tol {tol:.1e})", flush=True)
    assert err < tol, (what, err, tol)


def main():
    import torch
    import torch.distributed as dist

    from oracle import oracle as orc  # the checker (tests only)
    from Rigid import RigidBody
    from rigid_body_light_b200.sharding import PartitionedRigidBody
    from rigid_body_light_b200.shells import sphere_suspension

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count())
    dist.init_process_group("gloo", rank=rank, world_size=world)  # plumbing only: hands the NCCL id round
    n_bodies = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    s = sphere_suspension(n_bodies, 42, True)
    n_blb = 42
    n3, n6 = 3 * n_bodies * n_blb, 6 * n_bodies
    rng = np.random.default_rng(11)
    vec = rng.standard_normal(n3 + n6)
    rhs = np.concatenate([np.zeros(n3), rng.standard_normal(n6)])
    W = [rng.standard_normal(n3) for _ in range(3)]
    F_ext = np.tile(np.array([0, 0, -1.0, 0, 0, 0]), n_bodies)
    # tolerances ~10x the worst errors observed on B200 at world sizes 1, 2, 3, 4 and 8 (profiles/
    # r02_partitioned_tests_8gpu.log; every value is printed again below; a single context sums in another order)
    for precision, tol in (("double", 1e-12), ("single", 1e-5)):
        for block in (False, True):
            one = RigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, block_PC=block, precision=precision)
            part = PartitionedRigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, block_PC=block,
                                        precision=precision, rank=rank, world=world, dist=dist, force_comm=True)
            # the exchanges around every product: this library's peer-memory kernels where every rank could map
            # every other rank's buffer, NCCL collectives otherwise (include/rbl.h).  The whole battery below runs
            # with the default; the NCCL protocol is then checked against it on the saddle operator and a BD step.
            exch = part.exchange
            if rank == 0:
                print(f"OBS world={world} exchange {precision} {block}: {exch} {part.exchange_why!r}", flush=True)
            if os.environ.get("RBL_REQUIRE_PEER") == "1":
                assert exch == "peer", part.exchange_why
            # saddle operator: against the single context AND against the CPU oracle (on the positions the
            # ranks placed and the inputs rounded to the run's precision, so that <= 1e-5 / 1e-12 applies)
            got = part.gather_system(part.apply_saddle(part.slice_system(vec)))
            chk(rel(got, one.apply_saddle(vec)), tol, "saddle vs one context", precision, block)
            parts = [None] * world
            dist.all_gather_object(parts, part.get_blob_positions())
            ndt = np.float64 if precision == "double" else np.float32
            r_all = np.concatenate(parts).astype(np.float64)
            lam, U = vec[:n3].astype(ndt).astype(np.float64), vec[n3:].astype(ndt).astype(np.float64)
            Xp = s["X"].astype(ndt).astype(np.float64)
            want = np.concatenate([orc.apply_M(lam, r_all, s["a"], 1.0, True) - orc.K_dot(U, r_all, Xp, n_blb),
                                   orc.KT_dot(lam, r_all, Xp, n_blb)])
            chk(rel(got, want), 1e-12 if precision == "double" else 1e-5, "saddle vs ORACLE", precision, block)
            # preconditioner (rank-local)
            got = part.gather_system(part.apply_PC(part.slice_system(vec)))
            chk(rel(got, one.apply_PC(vec)), tol / 10, "pc vs one context", precision, block)  # rank-local: identical
            # GMRES
            gt = 1e-10 if precision == "double" else 1e-5
            x, it, rr = part.gmres(part.slice_system(rhs), tol=gt, restart=40, max_iter=120)
            x1, it1, rr1 = one.gmres(rhs, tol=gt, restart=40, max_iter=120)
            assert rr <= gt and abs(it - it1) <= 2, ("gmres", precision, block, it, it1, rr)
            chk(rel(part.gather_system(x), x1), 1e-12 if precision == "double" else 3e-5, "gmres x", precision, block)
            # Lanczos square root
            lt = 1e-9 if precision == "double" else 1e-5
            y, k = part.brownian_sqrt(part.slice_blobs(W[0]), tol=lt, max_iter=80)
            y1, k1 = one.brownian_sqrt(W[0], tol=lt, max_iter=80)
            parts = [None] * world
            dist.all_gather_object(parts, y)
            chk(rel(np.concatenate(parts), y1), 1e-13 if precision == "double" else 5e-6, "lanczos", precision, k, k1)
            # full Brownian step: same noise, same step
            noise_l = tuple(part.slice_blobs(w) for w in W)
            U, it, rr = part.bd_step(part.slice_bodies(F_ext), kBT=0.0041, noise_local=noise_l, tol=gt, restart=40,
                                     max_iter=120, lanczos_tol=lt, lanczos_max_iter=80)
            U1, it1, rr1 = one.bd_step(F_ext, kBT=0.0041, noise=tuple(W), tol=gt, restart=40, max_iter=120,
                                       lanczos_tol=lt, lanczos_max_iter=80)
            parts = [None] * world
            dist.all_gather_object(parts, U)
            chk(rel(np.concatenate(parts), U1), 1e-12 if precision == "double" else 2e-5, "bd_step U", precision, block)
            # device-generated noise: same (seed, step) on every rank = the single-context step
            U, it, rr = part.bd_step(part.slice_bodies(F_ext), kBT=0.0041, seed=1234, step=7, tol=gt, restart=40,
                                     max_iter=120, lanczos_tol=lt, lanczos_max_iter=80)
            U1, it1, rr1 = one.bd_step(F_ext, kBT=0.0041, seed=1234, step=7, tol=gt, restart=40, max_iter=120,
                                       lanczos_tol=lt, lanczos_max_iter=80)
            parts = [None] * world
            dist.all_gather_object(parts, U)
            chk(rel(np.concatenate(parts), U1), 1e-12 if precision == "double" else 2e-5, "seeded bd_step U", precision, block)
            Xp, Qp = part.get_config()
            X1, Q1 = one.get_config()
            chk(rel(Xp, X1[part.b0:part.b1]), 1e-14 if precision == "double" else 2e-7, "X after the step", precision, block)
            chk(rel(Qp, Q1[part.b0:part.b1]), 1e-14 if precision == "double" else 5e-7, "Q after the step", precision, block)
            if exch == "peer":
                # the same operators with the other protocol, switched at run time, at the configuration the step left
                one_now = one.apply_saddle(vec)
                for mode in ("nccl", "peer"):
                    part.set_exchange(mode)
                    assert part.exchange == mode
                    got = part.gather_system(part.apply_saddle(part.slice_system(vec)))
                    chk(rel(got, one_now), tol, f"saddle vs one context [{mode} after switch]", precision, block)
                part.set_exchange("nccl")
                U, it, rr = part.bd_step(part.slice_bodies(F_ext), kBT=0.0041, seed=99, step=3, tol=gt, restart=40,
                                         max_iter=120, lanczos_tol=lt, lanczos_max_iter=80)
                U1, it1, rr1 = one.bd_step(F_ext, kBT=0.0041, seed=99, step=3, tol=gt, restart=40, max_iter=120,
                                           lanczos_tol=lt, lanczos_max_iter=80)
                parts = [None] * world
                dist.all_gather_object(parts, U)
                chk(rel(np.concatenate(parts), U1), 1e-12 if precision == "double" else 2e-5, "seeded bd_step U [nccl]", precision, block)
                part.set_exchange("peer")
            if precision == "double":
                # mixed precision on the partitioned suspension (the float mirror shares the communicator and has
                # its own peer buffers): against the all-double single context, same (seed, step)
                for mode in (1, 2):
                    part.set_mixed_precision(mode)
                    U, it, rr = part.bd_step(part.slice_bodies(F_ext), kBT=0.0041, seed=77, step=mode, tol=gt, restart=40,
                                             max_iter=120, lanczos_tol=lt, lanczos_max_iter=80)
                    U1, it1, rr1 = one.bd_step(F_ext, kBT=0.0041, seed=77, step=mode, tol=gt, restart=40, max_iter=120,
                                               lanczos_tol=lt, lanczos_max_iter=80)
                    assert rr <= gt, ("mixed bd_step relres", mode, rr)
                    parts = [None] * world
                    dist.all_gather_object(parts, U)
                    # mode 1: same stopping rule on the double residual; mode 2: float-rounded operator under the root
                    chk(rel(np.concatenate(parts), U1), 2e-9 if mode == 1 else 3e-6, f"bd_step U [mixed {mode}]", precision, block)
                part.set_mixed_precision(0)
            part.close()
    # a blob below the wall on ONE rank must surface as the same error on EVERY rank
    Xb = s["X"].copy()
    Xb[0, 2] = 0.0
    part = PartitionedRigidBody(s["cfg"], Xb, s["Q"], s["a"], 1.0, 0.01, wall_PC=True, precision="double",
                                rank=rank, world=world, dist=dist, force_comm=True)
    try:
        part.apply_saddle(part.slice_system(vec))
        raise AssertionError("no error for a blob below the wall")
    except RuntimeError as exc:
        assert "BELOW_WALL" in str(exc) or "another rank" in str(exc), str(exc)
    part.close()
    dist.barrier()
    if rank == 0:
        print("PARTITIONED-OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
