"""Golden outputs of the reference's OWN random-finite-difference members (c_rigid_obj.cpp:712-728,
743-863, 880-893) with the noise injected, generated where /root/reference exists through
oracle.RefBody (the members compiled from the reference source, oracle/build_ref.sh):

    python tests/golden/make_rfd_golden.py        ->  tests/golden/rfd_ref_golden.npz

Cases: 6 touching spheres of shell_N_42 (a = sep/2), above the wall and in free space, float64 and
float32 builds of the reference.  tests/test_gpu_rfd.py compares the CUDA entry points with these."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from rigid_body_light_b200.shells import sphere_suspension  # noqa: E402


def main():
    out = {}
    rng = np.random.default_rng(20261018)
    nb, shell = 6, 42
    n3, n6 = 3 * nb * shell, 6 * nb
    W, U, W6 = rng.standard_normal(n3), rng.standard_normal(n6), rng.standard_normal(n6)
    out["W"], out["U"], out["W6"] = W, U, W6
    for wall in (True, False):
        s = sphere_suspension(nb, shell, wall)
        Q = s["Q"] * np.linspace(0.7, 1.4, nb)[:, None]  # un-normalised on purpose (setConfig normalises, :216)
        key = f"wall{int(wall)}"
        out[f"{key}/cfg"], out[f"{key}/X"], out[f"{key}/Q"], out[f"{key}/a"] = s["cfg"], s["X"], Q, s["a"]
        for dtype, tag in ((np.float64, "f64"), (np.float32, "f32")):
            rb = orc.RefBody(s["cfg"].astype(dtype), s["X"].astype(dtype), Q.astype(dtype), s["a"], 0.9, 0.02, wall_PC=wall, dtype=dtype)
            p = f"{key}/{tag}/"
            out[p + "M_RFD"] = rb.M_RFD(W)                      # delta 1e-4 (:771)
            out[p + "M_RFD_from_U"] = rb.M_RFD_from_U(U, W)     # delta 1e-3 (:820)
            out[p + "KT_RFD_from_U"] = rb.KT_RFD_from_U(U, W)   # delta 1e-3 (:844)
            out[p + "KTinv_RFD"] = rb.KTinv_RFD(W6)             # delta 1e-4 (:745)
            rp, rm = rb.M_RFD_cfgs(U, 1.0e-3)
            out[p + "r_plus"], out[p + "r_minus"] = rp, rm
            X, Qo = rb.update_X_Q_out(0.05 * U)
            out[p + "X_out"], out[p + "Q_out"] = X, Qo
            rb.evolve_RFD(0.05 * U)
            Xe, Qe = rb.get_config()
            out[p + "X_evolved"], out[p + "Q_evolved"], out[p + "r_evolved"] = Xe, Qe, rb.positions()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rfd_ref_golden.npz"), **out)
    print("wrote rfd_ref_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
