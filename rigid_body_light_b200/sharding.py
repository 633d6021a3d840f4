"""Multi-GPU sharding of the saddle operator: one process per GPU, bodies partitioned in
contiguous ranges.

Every rank owns the bodies [b0, b1): their blob placement, K, K^T and preconditioner are
rank-local.  The mobility product needs every source, so each application all-gathers the
constraint forces lambda (3N reals) over NCCL/NVLink; the O(N^2) work itself is the upper
triangle of the tile grid of the SYMMETRIC kernel, cut into ``world`` equal contiguous shares
(``rbl_dev_apply_M_part``): every rank produces a partial product over all blobs, the partial
products are summed with one reduce-scatter (NCCL, equal shards; an all-reduce otherwise) of 3N
reals, and each rank finishes its own rows with its local K and K^T (``rbl_dev_saddle_finish``).  Blob positions are all-gathered once per
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
        lo = 3 * self.t0
        if self.even and self.dist.get_backend() == "nccl":
            # every rank only needs the sum of ITS rows: reduce-scatter moves half of what an
            # all-reduce does
            if getattr(self, "_mine", None) is None:
                self._mine = self.torch.empty(n3, dtype=self._mbuf.dtype, device=self._mbuf.device)
            mine = self._mine
            self.dist.reduce_scatter_tensor(mine, self._mbuf)
        else:
            self.dist.all_reduce(self._mbuf)  # uneven shards / gloo: sum of the partial products everywhere
            mine = self._mbuf[lo:lo + n3]
        if ev:
            ev[3].record()
            self.timing.append(ev)
        self.backend.saddle_finish(mine, x_local[:n3], x_local[n3:], out_local)
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


def join_system(parts, ranges):
    """Inverse of ``slice_system``: the global [lambda ; U] vector from every rank's local one."""
    lam = [p[: p.size - 6 * (hi - lo)] for p, (lo, hi) in zip(parts, ranges)]
    U = [p[p.size - 6 * (hi - lo):] for p, (lo, hi) in zip(parts, ranges)]
    return np.concatenate(lam + U)


def slice_system(vec, ranges, n_blb, rank):
    """Rank-local [lambda ; U] slice of a global [lambda(3N) ; U(6 n_bod)] vector (numpy)."""
    n_bod = ranges[-1][1]
    n3 = 3 * n_bod * n_blb
    lo, hi = ranges[rank]
    return np.concatenate([vec[3 * lo * n_blb: 3 * hi * n_blb], vec[n3 + 6 * lo: n3 + 6 * hi]])


class PartitionedRigidBody:
    """One rank's handle on a suspension partitioned over the GPUs of a node, with NCCL inside
    the library (``rbl_comm_init``, include/rbl.h): the same operators as ``Rigid.RigidBody`` but
    on rank-local slices, and collective -- every rank calls them together.

    Every rank passes the GLOBAL ``X`` / ``Q``; the rank keeps the bodies ``ranges[rank]``.
    System vectors are ``[lambda_local (3 N_local) ; U_local (6 n_bod_local)]``
    (``slice_system`` cuts them out of a global vector, ``gather_system`` puts them back).

    ``dist`` is ``torch.distributed`` (any backend) and is used ONLY to hand rank 0's
    ncclUniqueId to the other ranks and by ``gather_system``; pass ``uid=`` to use another
    transport.  world_size 1 needs no communicator (``force_comm=True`` still creates one: the
    NCCL path with a single rank)."""

    def __init__(self, rigid_config, X, Q, a, eta, dt, wall_PC=False, block_PC=False, precision="single",
                 rank=0, world=1, dist=None, device=None, uid=None, force_comm=False):
        import ctypes

        from ._lib import Context

        rigid_config = np.asarray(rigid_config, dtype=np.float64).reshape(-1, 3)
        X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
        Q = np.asarray(Q, dtype=np.float64).reshape(-1, 4)
        if X.shape[0] != Q.shape[0]:
            raise RuntimeError("X and Q must have the same number of bodies")
        self.rank, self.world, self.dist = rank, world, dist
        self.n_blb = rigid_config.shape[0]
        self.N_bodies_global = X.shape[0]
        self.ranges = body_ranges(self.N_bodies_global, world)
        if min(hi - lo for lo, hi in self.ranges) < 1:
            raise RuntimeError("every rank needs at least one body")
        self.b0, self.b1 = self.ranges[rank]
        self.N_bodies = self.b1 - self.b0
        self.total_blobs = self.N_bodies * self.n_blb
        self.ctx = Context(precision, device=-1 if device is None else device)
        self.real = self.ctx.real
        self.ctx.set_parameters(a, dt, 1.0, eta, rigid_config)
        self.ctx.set_flags(int(block_PC), int(wall_PC))
        self.ctx.set_config(X[self.b0:self.b1], Q[self.b0:self.b1])
        if world == 1 and not force_comm:
            return  # a single context needs no communicator (force_comm: NCCL path with one rank, for tests)
        if uid is None:
            box = [None]
            if rank == 0:
                buf = ctypes.create_string_buffer(128)
                st = self.ctx.L.rbl_comm_unique_id(buf)
                if st != 0:
                    raise RuntimeError(self.ctx.L.rbl_last_error(None).decode())
                box[0] = buf.raw
            if world > 1:
                dist.broadcast_object_list(box, src=0)
            uid = box[0]
        counts = (ctypes.c_int * world)(*[(hi - lo) * self.n_blb for lo, hi in self.ranges])
        self.ctx.call("rbl_comm_init", ctypes.c_char_p(uid), rank, world, counts)

    # -- the exchanges around every product (include/rbl.h: peer memory, or NCCL collectives) ------
    @property
    def exchange(self):
        """'peer' (this library's kernels over NVLink peer memory), 'nccl', or 'none' (no communicator)."""
        if self.ctx.L.rbl_comm_exchange(self.ctx.h):
            return "peer"
        return "none" if self.ctx.L.rbl_comm_exchange_why(self.ctx.h) == b"no communicator" else "nccl"

    @property
    def exchange_why(self):
        return (self.ctx.L.rbl_comm_exchange_why(self.ctx.h) or b"").decode()

    def set_exchange(self, mode):
        """Collective: 'peer' or 'nccl' on every rank."""
        self.ctx.call("rbl_comm_set_exchange", {"nccl": 0, "peer": 1}[mode])

    def apply_saddle_dev(self, x_ptr, out_ptr):
        """Device pointers, asynchronous on the context's stream (``rbl_dev_apply_saddle``); collective."""
        self.ctx.call("rbl_dev_apply_saddle", x_ptr, out_ptr)

    # -- helpers ---------------------------------------------------------------------------
    def _in(self, v, n, what):
        v = np.ascontiguousarray(np.asarray(v, dtype=self.real).reshape(-1))
        if v.size != n:
            raise RuntimeError(f"{what} must have total size {n} on this rank, got {v.size}")
        return v

    @property
    def sys_size(self):
        return 3 * self.total_blobs + 6 * self.N_bodies

    def slice_system(self, vec_global):
        return slice_system(np.asarray(vec_global).reshape(-1), self.ranges, self.n_blb, self.rank)

    def slice_blobs(self, vec_global):
        v = np.asarray(vec_global).reshape(-1)
        return v[3 * self.b0 * self.n_blb: 3 * self.b1 * self.n_blb]

    def slice_bodies(self, vec_global, per=6):
        v = np.asarray(vec_global).reshape(-1)
        return v[per * self.b0: per * self.b1]

    def gather_system(self, x_local):
        """Global [lambda ; U] vector on every rank (host side, for checks and output)."""
        if self.world == 1:
            return np.asarray(x_local).copy()
        parts = [None] * self.world
        self.dist.all_gather_object(parts, np.asarray(x_local))
        return join_system(parts, self.ranges)

    def set_mixed_precision(self, mode):
        """see Rigid.RigidBody.set_mixed_precision (double contexts; on a partitioned suspension the float mirror
        shares the communicator -- collective, every rank the same mode)"""
        self.ctx.call("rbl_set_mixed_precision", int(mode))

    def get_blob_positions(self):
        """Positions of THIS rank's blobs, (N_local, 3) (multi_body_pos, c_rigid_obj.cpp:295-300)."""
        r = np.empty(3 * self.total_blobs, dtype=self.real)
        self.ctx.call("rbl_blob_positions", r.ctypes.data)
        return r.reshape(-1, 3)

    # -- collective operators (rank-local slices in and out) ---------------------------------
    def apply_saddle(self, x_local):
        x = self._in(x_local, self.sys_size, "x")
        out = np.empty_like(x)
        self.ctx.call("rbl_apply_saddle", x.ctypes.data, out.ctypes.data)
        return out

    def apply_PC(self, b_local):  # rank-local: whole bodies per rank
        b = self._in(b_local, self.sys_size, "b")
        out = np.empty_like(b)
        self.ctx.call("rbl_apply_PC", b.ctypes.data, out.ctypes.data)
        return out

    def gmres(self, rhs_local, tol=1e-8, restart=60, max_iter=300):
        import ctypes

        rhs = self._in(rhs_local, self.sys_size, "rhs")
        x = np.empty_like(rhs)
        it, rr = ctypes.c_int(), ctypes.c_double()
        self.ctx.call("rbl_gmres", rhs.ctypes.data, x.ctypes.data, tol, restart, max_iter, ctypes.byref(it), ctypes.byref(rr))
        return x, it.value, rr.value

    def brownian_sqrt(self, W_local, tol=1e-6, max_iter=100):
        import ctypes

        W = self._in(W_local, 3 * self.total_blobs, "W")
        out = np.empty_like(W)
        it = ctypes.c_int()
        self.ctx.call("rbl_lanczos_sqrt", W.ctypes.data, out.ctypes.data, tol, max_iter, ctypes.byref(it))
        return out, it.value

    def bd_step(self, F_ext_local, slip_local=None, kBT=0.0, noise_local=None, tol=1e-8, restart=60, max_iter=300,
                lanczos_tol=1e-6, lanczos_max_iter=100, seed=None, step=0):
        """One BD step of the whole suspension (``rbl_bd_step``); the rank's bodies move.
        ``noise_local`` = (W1, W2, Wr) slices of three GLOBAL standard-normal vectors, or ``seed`` /
        ``step`` for the device generator (``rbl_bd_step_seeded``): every rank passes the same seed
        and step, each element's noise depends on its GLOBAL index only."""
        import ctypes

        n3 = 3 * self.total_blobs
        F = self._in(F_ext_local, 6 * self.N_bodies, "F_ext")
        slip = None if slip_local is None else self._in(slip_local, n3, "slip")
        if seed is not None and kBT > 0 and noise_local is None:
            U = np.empty(6 * self.N_bodies, dtype=self.real)
            it, rr = ctypes.c_int(), ctypes.c_double()
            self.ctx.call("rbl_bd_step_seeded", F.ctypes.data, None if slip is None else slip.ctypes.data, int(seed), int(step),
                          float(kBT), tol, restart, max_iter, lanczos_tol, lanczos_max_iter, U.ctypes.data, ctypes.byref(it),
                          ctypes.byref(rr))
            return U, it.value, rr.value
        W = [None, None, None]
        if kBT > 0:
            if noise_local is None:
                raise RuntimeError("kBT > 0 needs noise_local = (W1, W2, Wr)")
            W = [self._in(w, n3, "noise") for w in noise_local]
        U = np.empty(6 * self.N_bodies, dtype=self.real)
        it, rr = ctypes.c_int(), ctypes.c_double()
        ptr = lambda v: None if v is None else v.ctypes.data  # noqa: E731
        self.ctx.call("rbl_bd_step", F.ctypes.data, ptr(slip), ptr(W[0]), ptr(W[1]), ptr(W[2]), float(kBT), tol, restart,
                      max_iter, lanczos_tol, lanczos_max_iter, U.ctypes.data, ctypes.byref(it), ctypes.byref(rr))
        return U, it.value, rr.value

    def get_config(self):
        X = np.empty(3 * self.N_bodies, dtype=self.real)
        Q = np.empty(4 * self.N_bodies, dtype=self.real)
        self.ctx.call("rbl_get_config", X.ctypes.data, Q.ctypes.data)
        return X.reshape(-1, 3), Q.reshape(-1, 4)

    def close(self):
        self.ctx.close()
