"""The N>1 path on CPU: two (and three, uneven) gloo ranks run ShardedSaddle with an
oracle-backed arithmetic backend injected; the concatenated rank outputs must equal the
single-process saddle operator.  Exercises body partitioning, the padded all-gather for
uneven shards and the global/local index bookkeeping (the GPU arithmetic itself is covered
by tests/test_gpu_matvec.py::test_sharded_target_ranges_tile_the_full_product)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_bodies, shell, out_q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import oracle as orc
    from rigid_body_light_b200.sharding import ShardedSaddle, body_ranges, slice_system
    from rigid_body_light_b200.shells import sphere_suspension

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = sphere_suspension(n_bodies, shell, True)
    ref = orc.remove_mean(s["cfg"])
    n_blb = ref.shape[0]
    ranges = body_ranges(n_bodies, world)
    lo, hi = ranges[rank]
    Xl, Ql = s["X"][lo:hi], s["Q"][lo:hi]

    class OracleShard:  # test double for CudaShard: same interface, CPU oracle arithmetic
        def positions(self):
            return torch.from_numpy(orc.blob_positions(Xl, Ql, ref).reshape(-1).copy())

        def saddle_shard(self, lam_all, r_all, n_all, t0, U_local, out_local):
            nl = (hi - lo) * n_blb
            rl = r_all.numpy().reshape(-1, 3)[t0:t0 + nl]
            slip = orc.apply_M(lam_all.numpy(), r_all.numpy(), s["a"], 1.0, True, rows=(t0, nl))
            slip = slip - orc.K_dot(U_local.numpy(), rl, Xl, n_blb)
            F = orc.KT_dot(lam_all.numpy()[3 * t0:3 * (t0 + nl)], rl, Xl, n_blb)
            out_local.copy_(torch.from_numpy(np.concatenate([slip, F])))

        def apply_M_part(self, lam_all, r_all, n_all, part, n_parts, out_all):
            # any decomposition whose parts SUM to the product will do for the host logic: this
            # double gives part p the rows [n p/P, n (p+1)/P) and zeros elsewhere
            r0, r1 = n_all * part // n_parts, n_all * (part + 1) // n_parts
            out = np.zeros(3 * n_all)
            out[3 * r0:3 * r1] = orc.apply_M(lam_all.numpy(), r_all.numpy(), s["a"], 1.0, True, rows=(r0, r1 - r0))
            out_all.copy_(torch.from_numpy(out))

        def saddle_finish(self, Mlam_local, lam_local, U_local, out_local):
            nl = (hi - lo) * n_blb
            rl = orc.blob_positions(Xl, Ql, ref)
            slip = Mlam_local.numpy() - orc.K_dot(U_local.numpy(), rl, Xl, n_blb)
            F = orc.KT_dot(lam_local.numpy(), rl, Xl, n_blb)
            assert slip.size == 3 * nl
            out_local.copy_(torch.from_numpy(np.concatenate([slip, F])))

    op = ShardedSaddle(OracleShard(), n_bodies, n_blb, rank, world, dist)
    vec = np.random.default_rng(2).standard_normal(3 * n_bodies * n_blb + 6 * n_bodies)
    x_local = torch.from_numpy(slice_system(vec, ranges, n_blb, rank))
    out = torch.empty_like(x_local)
    op.apply(x_local, out)
    out_q.put((rank, out.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_bodies", [(2, 6), (3, 7)])
def test_sharded_saddle_matches_single_process(orc, world, n_bodies):
    import torch.multiprocessing as mp

    from rigid_body_light_b200.sharding import body_ranges
    from rigid_body_light_b200.shells import sphere_suspension

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_bodies, 12, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = sphere_suspension(n_bodies, 12, True)
    ref = orc.remove_mean(s["cfg"])
    vec = np.random.default_rng(2).standard_normal(3 * n_bodies * 12 + 6 * n_bodies)
    want = orc.apply_saddle(vec, s["X"], s["Q"], ref, s["a"], 1.0, True)
    n3 = 3 * n_bodies * 12
    ranges = body_ranges(n_bodies, world)
    slip = np.concatenate([results[r][: 3 * (hi - lo) * 12] for r, (lo, hi) in enumerate(ranges)])
    F = np.concatenate([results[r][3 * (hi - lo) * 12:] for r, (lo, hi) in enumerate(ranges)])
    assert np.allclose(slip, want[:n3], rtol=1e-13, atol=1e-14)
    assert np.allclose(F, want[n3:], rtol=1e-13, atol=1e-14)


def test_body_ranges_are_contiguous_and_balanced():
    from rigid_body_light_b200.sharding import body_ranges

    for n, w in [(1000, 8), (1000, 3), (7, 3), (4096, 8), (5, 8)]:
        r = body_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1


def test_slice_and_join_system_are_inverse_for_uneven_ranges():
    from rigid_body_light_b200.sharding import body_ranges, join_system, slice_system

    for n_bod, world, n_blb in [(7, 3, 12), (1000, 8, 5), (5, 2, 42), (4, 1, 3)]:
        ranges = body_ranges(n_bod, world)
        vec = np.arange(3 * n_bod * n_blb + 6 * n_bod, dtype=np.float64)
        parts = [slice_system(vec, ranges, n_blb, r) for r in range(world)]
        assert [p.size for p in parts] == [(hi - lo) * (3 * n_blb + 6) for lo, hi in ranges]
        assert np.array_equal(join_system(parts, ranges), vec)
