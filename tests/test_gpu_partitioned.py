"""Partitioned suspensions (NCCL inside librbl, include/rbl.h `rbl_comm_init`): the collective
saddle operator, GMRES, Lanczos and the BD step on rank-local slices must agree with one
context holding the whole suspension.  World size 1 runs on any GPU box (the full NCCL code
path with a single rank); world sizes 2, 3, 4 and 8 (uneven body ranges) need that many GPUs.  The
worker also compares the partitioned saddle operator with the CPU oracle, and prints every observed error."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
WORKER = os.path.join(ROOT, "tests", "host", "partitioned_worker.py")


def _run(world, n_bodies):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29000 + os.getpid() % 2000 + world), WORKER, str(n_bodies)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0 and "PARTITIONED-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    obs = [ln for ln in r.stdout.splitlines() if ln.startswith("OBS ")]
    print("\n".join(obs))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"partitioned_world{world}.log"), "w") as fh:
            fh.write("\n".join(obs) + "\nPARTITIONED-OK\n")


def test_partitioned_world1_matches_single_context():
    _run(1, 4)


@pytest.mark.parametrize("world,n_bodies", [(2, 5), (3, 7), (4, 9), (8, 19)])
def test_partitioned_multi_gpu_matches_single_context(world, n_bodies):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(world, n_bodies)


def test_exchange_entry_points_without_a_communicator():
    """A context that never called rbl_comm_init reports no exchange and refuses to switch."""
    from rigid_body_light_b200._lib import Context, RblError

    ctx = Context("single")
    assert ctx.L.rbl_comm_exchange(ctx.h) == 0
    assert ctx.L.rbl_comm_exchange_why(ctx.h) == b"no communicator"
    assert ctx.L.rbl_comm_world(ctx.h) == 1
    with pytest.raises(RblError, match="no communicator"):
        ctx.call("rbl_comm_set_exchange", 1)
    ctx.close()
