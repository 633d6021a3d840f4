"""Multi-GPU sharding of the saddle operator: one process per GPU, bodies partitioned in
contiguous ranges.

Every rank owns the bodies [b0, b1): their blob placement, K, K^T and preconditioner are
rank-local.  The mobility product needs every source, so each application all-gathers the
constraint forces lambda (3N reals) over NCCL/NVLink; the O(N^2) work itself is the upper
triangle of the tile grid of the SYMMETRIC kernel, cut into ``world`` equal contiguous shares
(``rbl_dev_apply_M_part``): every rank produces a partial product over all blobs, the partial
products are summed with one all-reduce (3N reals), and each rank finishes its own rows with
its local K and K^T (``rbl_dev_saddle_finish``).  Blob positions are all-gathered once per
configuration, not per application.

The collective plumbing is ``torch.distributed`` (backend "nccl" on GPUs; the same code runs
under "gloo" on CPU tensors for the host-logic tests).  The arithmetic is behind a small
backend interface: ``CudaShard`` (the product; C ABI, no CPU fallback) or whatever a test
injects.
"""
from __future__ import annotations

import numpy as np


def body_ranges(n_bodies: int, world: int):
    """Contiguous, balanced body ranges: the first (n_bodies % world) ranks get one more."""
    base, extra = divmod(n_bodies, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


class ShardedSaddle:
    """apply_saddle over ``world`` ranks.  ``backend`` must provide
    ``positions() -> tensor(3*n_local)``,
    ``saddle_shard(lam_all, r_all, n_all, t0, U_local, out_local)`` (single rank),
    ``apply_M_part(lam_all, r_all, n_all, part, n_parts, out_all)`` and
    ``saddle_finish(Mlam_local, lam_local, U_local, out_local)``."""

    def __init__(self, backend, n_bodies, n_blb, rank, world, dist=None):
        import torch

        self.torch = torch
        self.dist = dist
        self.backend = backend
        self.rank, self.world = rank, world
        self.n_blb = n_blb
        self.ranges = body_ranges(n_bodies, world)
        self.b0, self.b1 = self.ranges[rank]
        self.n_all = n_bodies * n_blb
        self.t0 = self.b0 * n_blb
        self.n_local = (self.b1 - self.b0) * n_blb
        # all_gather_into_tensor needs equal contributions: pad every rank to the largest shard
        self.max_local = max(hi - lo for lo, hi in self.ranges) * n_blb
        self.even = all((hi - lo) * n_blb == self.max_local for lo, hi in self.ranges)
        self.r_all = None
        self._gather_buf = None
        self._pad = None
        self.timing = None  # set to [] to record (allgather, product, allreduce) CUDA-event triples per apply

    def _allgather(self, local, out_all):
        """out_all (3*n_all) <- concatenation of every rank's 3*n_local slice."""
        torch, dist = self.torch, self.dist
        if self.world == 1:
            out_all.copy_(local)
            return
        if self.even:
            dist.all_gather_into_tensor(out_all, local)
            return
        if self._gather_buf is None:
            self._gather_buf = torch.empty(self.world * 3 * self.max_local, dtype=local.dtype, device=local.device)
            self._pad = torch.zeros(3 * self.max_local, dtype=local.dtype, device=local.device)
        self._pad[: local.numel()].copy_(local)
        dist.all_gather_into_tensor(self._gather_buf, self._pad)
        for r, (lo, hi) in enumerate(self.ranges):
            n = 3 * (hi - lo) * self.n_blb
            out_all[3 * lo * self.n_blb: 3 * lo * self.n_blb + n].copy_(
                self._gather_buf[r * 3 * self.max_local: r * 3 * self.max_local + n])

    def refresh_positions(self):
        """All-gather blob positions (once per configuration change)."""
        local = self.backend.positions()
        if self.r_all is None:
            self.r_all = self.torch.empty(3 * self.n_all, dtype=local.dtype, device=local.device)
            self.lam_all = self.torch.empty_like(self.r_all)
        self._allgather(local, self.r_all)

    def apply(self, x_local, out_local):
        """x_local = [lambda_local (3 n_local) ; U_local (6 nb_local)] -> out_local, same layout."""
        if self.r_all is None:
            self.refresh_positions()
        n3 = 3 * self.n_local
        ev = None
        if self.timing is not None and self.world > 1:
            ev = [self.torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        self._allgather(x_local[:n3], self.lam_all)
        if ev:
            ev[1].record()
        if self.world == 1:
            self.backend.saddle_shard(self.lam_all, self.r_all, self.n_all, self.t0, x_local[n3:], out_local)
            return out_local
        if getattr(self, "_mbuf", None) is None:
            self._mbuf = self.torch.empty_like(self.lam_all)
        self.backend.apply_M_part(self.lam_all, self.r_all, self.n_all, self.rank, self.world, self._mbuf)
        if ev:
            ev[2].record()
        self.dist.all_reduce(self._mbuf)  # sum of the partial products
        if ev:
            ev[3].record()
            self.timing.append(ev)
        lo = 3 * self.t0
        self.backend.saddle_finish(self._mbuf[lo:lo + n3], x_local[:n3], x_local[n3:], out_local)
        return out_local


def comm_breakdown(timing):
    """(allgather_ms, product_ms, allreduce_ms) averaged over the recorded applies."""
    if not timing:
        return None
    ag = sum(e[0].elapsed_time(e[1]) for e in timing) / len(timing)
    pr = sum(e[1].elapsed_time(e[2]) for e in timing) / len(timing)
    ar = sum(e[2].elapsed_time(e[3]) for e in timing) / len(timing)
    return ag, pr, ar


class CudaShard:
    """Rank-local arithmetic on the GPU through the C ABI (include/rbl.h)."""

    def __init__(self, ctx, n_local_blobs, torch_dtype):
        import torch

        self.ctx = ctx
        self.torch = torch
        self.n_local = n_local_blobs
        self.dtype = torch_dtype
        # run on torch's current stream so NCCL collectives and kernels are ordered by torch
        ctx.call("rbl_set_stream", torch.cuda.current_stream().cuda_stream)

    def positions(self):
        r = self.torch.empty(3 * self.n_local, dtype=self.dtype, device="cuda")
        self.ctx.call("rbl_dev_blob_positions", r.data_ptr())
        return r

    def saddle_shard(self, lam_all, r_all, n_all, t0, U_local, out_local):
        self.ctx.call("rbl_dev_apply_saddle_shard", lam_all.data_ptr(), r_all.data_ptr(), n_all, t0,
                      U_local.data_ptr(), out_local.data_ptr())

    def apply_M_part(self, lam_all, r_all, n_all, part, n_parts, out_all):
        self.ctx.call("rbl_dev_apply_M_part", lam_all.data_ptr(), r_all.data_ptr(), n_all, part, n_parts,
                      out_all.data_ptr())

    def saddle_finish(self, Mlam_local, lam_local, U_local, out_local):
        self.ctx.call("rbl_dev_saddle_finish", Mlam_local.data_ptr(), lam_local.data_ptr(), U_local.data_ptr(),
                      out_local.data_ptr())


def slice_system(vec, ranges, n_blb, rank):
    """Rank-local [lambda ; U] slice of a global [lambda(3N) ; U(6 n_bod)] vector (numpy)."""
    n_bod = ranges[-1][1]
    n3 = 3 * n_bod * n_blb
    lo, hi = ranges[rank]
    return np.concatenate([vec[3 * lo * n_blb: 3 * hi * n_blb], vec[n3 + 6 * lo: n3 + 6 * hi]])
