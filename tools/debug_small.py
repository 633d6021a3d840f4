import sys, os, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rigid_body_light_b200._lib import Context
from rigid_body_light_b200.shells import sphere_suspension
from rigid_body_light_b200.sharding import CudaShard, ShardedSaddle
s = sphere_suspension(64, 42, True); ref = s["cfg"] - s["cfg"].mean(axis=0)
nb, n_blb = 64, 42; n = nb * n_blb
vec = np.random.default_rng(2).standard_normal(3 * n + 6 * nb)
for precision, tdt, ndt in (("single", torch.float32, np.float32), ("double", torch.float64, np.float64)):
    ctx = Context(precision)
    ctx.set_parameters(s["a"], 0.01, 1.0, 1.0, ref); ctx.set_flags(0, 1); ctx.set_config(s["X"], s["Q"])
    x = vec.astype(ndt); out_h = np.empty_like(x)
    ctx.call("rbl_apply_saddle", x.ctypes.data, out_h.ctypes.data)
    out_h2 = np.empty_like(x)
    ctx.call("rbl_apply_saddle", x.ctypes.data, out_h2.ctypes.data)
    shard = CudaShard(ctx, n, tdt)
    op = ShardedSaddle(shard, nb, n_blb, 0, 1, None)
    xd = torch.from_numpy(x).cuda(); od = torch.empty_like(xd)
    op.apply(xd, od); torch.cuda.synchronize()
    o1 = od.cpu().numpy()
    ctx.call("rbl_flush_l2"); op.apply(xd, od); torch.cuda.synchronize(); o2 = od.cpu().numpy()
    out_h3 = np.empty_like(x)
    ctx.call("rbl_apply_saddle", x.ctypes.data, out_h3.ctypes.data)
    d = np.abs(o1 - out_h)
    print(precision, "host==host2", np.array_equal(out_h, out_h2), "dev1==dev2", np.array_equal(o1, o2), "dev==host", np.array_equal(o1, out_h),
          "host3==host", np.array_equal(out_h3, out_h), "maxdiff", d.max(), "argmax", d.argmax(), "n3", 3 * n, "ndiff", (d > 0).sum())
    r_all = op.r_all.cpu().numpy(); rh = np.empty(3 * n, ndt); ctx.call("rbl_blob_positions", rh.ctypes.data)
    print("  positions equal", np.array_equal(r_all, rh))
