"""Peer-memory exchange against the NCCL collectives on the SAME contexts, alternating blocks of timed steps
(torchrun, N >= 2).  Prints one JSON line per block on rank 0: which exchange, ms per step (max over ranks)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from bench import build_suspension
    from rigid_body_light_b200.sharding import CudaShard, PartitionedRigidBody, body_ranges, slice_system

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    precision = sys.argv[1] if len(sys.argv) > 1 else "single"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    s = build_suspension("cfg2")
    nb, n_blb = s["n_bodies"], s["n_blb"]
    ranges = body_ranges(nb, world)
    lo, hi = ranges[rank]
    pb = PartitionedRigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, precision=precision,
                              rank=rank, world=world, dist=dist, device=local)
    ctx = pb.ctx
    tdt = torch.float32 if precision == "single" else torch.float64
    CudaShard(ctx, (hi - lo) * n_blb, tdt)  # context on torch's current stream
    vec = np.random.default_rng(2).standard_normal(3 * nb * n_blb + 6 * nb)
    x = torch.from_numpy(slice_system(vec, ranges, n_blb, rank).astype(np.float32 if precision == "single" else np.float64)).cuda()
    out = torch.empty_like(x)

    def step():
        ctx.call("rbl_flush_l2")
        pb.apply_saddle_dev(x.data_ptr(), out.data_ptr())

    def block(mode):
        pb.set_exchange(mode)
        for _ in range(3):
            step()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"exchange": mode, "precision": precision, "n_gpus": world, "ms_per_step": float(t.item())}), flush=True)

    assert pb.exchange == "peer", pb.exchange_why
    for mode in ("peer", "nccl", "peer", "nccl", "nccl", "peer", "nccl", "peer"):
        block(mode)
    dist.barrier()
    pb.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
