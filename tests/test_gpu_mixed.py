"""Mixed precision for double contexts (include/rbl.h rbl_set_mixed_precision): float GMRES corrections
inside an iterative refinement on the DOUBLE residual must deliver the double answer; float products
inside the Lanczos noise must deliver the square root of an operator that is M to float rounding."""
import numpy as np
import pytest

from conftest import check, load_golden, rel_err
from test_gpu_rigid import _solver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("block", [False, True])
@pytest.mark.parametrize("name", ["case_touch_wall", "case_touch_free", "case_overlap_wall"])
def test_mixed_gmres_reaches_the_double_solution(orc, name, block):
    g = load_golden(name)
    rhs = g["vec"]
    ref = _solver(g, "double", block=block)
    x64, it64, rr64 = ref.gmres(rhs, tol=1e-10, restart=80, max_iter=400)
    cb = _solver(g, "double", block=block)
    cb.set_mixed_precision(1)
    x, it, rr = cb.gmres(rhs, tol=1e-10, restart=80, max_iter=400)
    assert rr <= 1e-10, rr
    check(rel_err(ref.apply_saddle(x), rhs), 2e-10, "true DOUBLE residual of the mixed-precision solution (requested 1e-10)")
    a, eta, wall = float(g["a"]), float(g["eta"]), bool(g["wall"])
    M = np.asarray(orc.dense_mobility(g["r"], a, eta, wall))
    K = orc.K_dense(g["r"], g["X"], g["cfg"].shape[0])
    A = np.block([[M, -K], [K.T, np.zeros((K.shape[1], K.shape[1]))]])
    cond = np.linalg.cond(A)
    print(f"[{name}, block={block}] float iterations {it} (double solver: {it64}), cond(A) = {cond:.1e}")
    check(rel_err(x, np.linalg.solve(A, rhs)), 1e-9, f"mixed GMRES solution vs dense solve, cond {cond:.1e} x relres 1e-10")
    check(rel_err(x, x64), 1e-9, "mixed vs double GMRES solution")


@pytest.mark.parametrize("name", ["case_touch_wall", "case_touch_free"])
def test_mixed_bd_step(name):
    """Same noise: mode 1 reproduces the double step's U to 1e-8 (the solve is refined to the double
    residual); mode 2 changes the Brownian increments at float rounding (1e-7 of the operator)."""
    g = load_golden(name)
    nb, n3 = g["X"].shape[0], g["r"].size
    rng = np.random.default_rng(31)
    F = rng.standard_normal(6 * nb)
    noise = tuple(rng.standard_normal(n3) for _ in range(3))
    kw = dict(kBT=0.004, noise=noise, tol=1e-10, restart=100, max_iter=400, lanczos_tol=1e-9, lanczos_max_iter=200)
    out = {}
    for mode in (0, 1, 2):
        cb = _solver(g, "double", block=True)
        cb.set_mixed_precision(mode)
        U, it, rr = cb.bd_step(F, **kw)
        assert rr <= 1e-10, (mode, rr)
        out[mode] = (U, cb.get_config())
    check(rel_err(out[1][0], out[0][0]), 2e-9, "BD step U, mixed mode 1 vs all-double")  # observed 1.4e-10
    check(rel_err(out[1][1][0], out[0][1][0]), 1e-10, "X after the step, mixed mode 1 vs all-double")
    e2 = rel_err(out[2][0], out[0][0])
    print(f"[{name}] mode 2 (float products inside Lanczos): U differs from the all-double step by {e2:.2e}")
    check(e2, 3e-6, "BD step U, mixed mode 2 vs all-double: float-rounded operator under the square root")


def test_mixed_mode_is_refused_for_float_contexts():
    g = load_golden("case_touch_free")
    cb = _solver(g, "single")
    with pytest.raises(RuntimeError):
        cb.set_mixed_precision(1)
    cb.set_mixed_precision(0)
