"""Parity of the CUDA mobility product with the oracle (GPU box only, -m gpu).

Calls go through the reference-facing host class (Rigid.RigidBody -> c_rigid.CManyBodies ->
C ABI) and, for the sharded/device entry points, through the C ABI directly via ctypes.
Tolerances are BASELINE.json's: <= 1e-12 relative L2 in double, <= 1e-5 in float, with the
float comparison made on the SAME float32-representable inputs on both sides."""
import ctypes

import numpy as np
import pytest

from conftest import CASE_NAMES, TOL, load_golden, rel_err, check

pytestmark = pytest.mark.gpu

PRECISIONS = ["double", "single"]


def _dtype(precision):
    return np.float64 if precision == "double" else np.float32


def _solver(g, precision, wall=None, block=False):
    from Rigid import RigidBody

    wall = bool(g["wall"]) if wall is None else wall
    return RigidBody(g["cfg"], g["X"], g["Q"], float(g["a"]), float(g["eta"]), float(g["dt"]),
                     wall_PC=wall, block_PC=block, precision=precision)


def _oracle_for(orc, F, r, a, eta, wall, precision, rows=None):
    dt = _dtype(precision)
    F = np.asarray(F, dt).astype(np.float64)
    r = np.asarray(r, dt).astype(np.float64)
    return orc.apply_M(F, r, a, eta, wall, rows=rows)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_apply_M_matches_golden(orc, name, precision):
    g = load_golden(name)
    cb = _solver(g, precision)
    assert cb.precision == precision
    out = cb.apply_M(g["lam"], g["r"])
    assert out.dtype == _dtype(precision) and out.shape == (g["lam"].size,)
    if precision == "double":
        check(rel_err(out, g["MF"]), TOL["double"])
    else:
        want = _oracle_for(orc, g["lam"], g["r"], float(g["a"]), float(g["eta"]), bool(g["wall"]), precision)
        check(rel_err(out, want), TOL["single"])


@pytest.mark.parametrize("precision", PRECISIONS)
def test_apply_M_extra_free_blob(orc, precision):
    """positions/forces may be longer than N_bodies*N_blobs (tests/test_interface.py:171-177)."""
    g = load_golden("case_overlap_free")
    cb = _solver(g, precision)
    r = np.concatenate([g["r"].reshape(-1), [7.0, -3.0, 2.0]])
    F = np.concatenate([g["lam"], [0.3, -0.2, 0.9]])
    out = cb.apply_M(F, r)
    check(rel_err(out, _oracle_for(orc, F, r, float(g["a"]), float(g["eta"]), False, precision)), TOL[precision])


def _random_cloud(n, wall, seed, a=0.11):
    """non-overlapping-ish random blobs with a few deliberately overlapping pairs"""
    rng = np.random.default_rng(seed)
    side = max(2.0, (n ** (1 / 3)) * 3 * a)
    r = rng.uniform(0, side, (n, 3))
    if wall:
        r[:, 2] = rng.uniform(0.2 * a, side, n)  # some inside the damping layer z < a
    if n > 4:
        r[1] = r[0] + [0.7 * a, 0, 0.1 * a]  # overlapping pair: r < 2a branch
        r[3] = r[2] + [0, 2.0 * a, 0]        # exactly touching
    return r, rng.standard_normal(3 * n)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("wall", [False, True])
@pytest.mark.parametrize("n", [1, 2, 31, 255, 256, 257, 1000, 1025, 4099])
def test_ragged_sizes(orc, n, wall, precision):
    """empty-ish / ragged inputs around the 256-source tile and the target-tile sizes"""
    from rigid_body_light_b200._lib import Context

    a, eta = 0.11, 0.9
    r, F = _random_cloud(n, wall, seed=n)
    ctx = Context(precision)
    ctx.set_parameters(a, 0.01, 1.0, eta, np.zeros((1, 3)))
    ctx.set_flags(0, wall)
    out = ctx.apply_M(F, r)
    check(rel_err(out, _oracle_for(orc, F, r, a, eta, wall, precision)), TOL[precision])
    ctx.close()


def test_zero_blobs_is_a_no_op():
    from rigid_body_light_b200._lib import Context

    ctx = Context("double")
    ctx.set_parameters(0.1, 0.01, 1.0, 1.0, np.zeros((1, 3)))
    assert ctx.apply_M(np.zeros(0), np.zeros(0)).size == 0
    ctx.close()


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("wall", [False, True])
def test_every_kernel_variant_agrees(orc, wall, precision):
    """all (targets/thread x threads/CTA) tile shapes compute the same product"""
    from rigid_body_light_b200._lib import Context

    a, eta = 0.11, 1.0
    r, F = _random_cloud(3000, wall, seed=99)
    want = _oracle_for(orc, F, r, a, eta, wall, precision)
    ctx = Context(precision)
    ctx.set_parameters(a, 0.01, 1.0, eta, np.zeros((1, 3)))
    ctx.set_flags(0, wall)
    nv = ctx.L.rbl_num_matvec_variants(ctx.h)
    assert nv >= 2
    ctx.call("rbl_set_matvec_mode", 1)  # ordered kernel
    for v in range(nv):
        ctx.call("rbl_set_matvec_variant", v)
        out = ctx.apply_M(F, r)
        check(rel_err(out, want), TOL[precision])
        assert np.array_equal(ctx.apply_M(F, r), out), v  # bit-reproducible
    ctx.call("rbl_set_matvec_mode", 0)  # symmetric kernel
    ns = ctx.L.rbl_num_sym_variants(ctx.h)
    assert ns >= 2
    for v in range(ns):
        ctx.call("rbl_set_sym_variant", v)
        check(rel_err(ctx.apply_M(F, r), want), TOL[precision])
    ctx.close()


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("mode", [0, 1])
def test_both_kernels_on_a_suspension(orc, mode, precision):
    """ordered and symmetric kernels against the oracle on touching spheres above the wall
    (several target tiles, several source tiles, far and near tile pairs)"""
    from rigid_body_light_b200._lib import Context
    from rigid_body_light_b200.shells import sphere_suspension

    s = sphere_suspension(30, 162, True)
    ref = s["cfg"] - s["cfg"].mean(axis=0)
    ctx = Context(precision)
    ctx.set_parameters(s["a"], 0.01, 1.0, 1.0, ref)
    ctx.set_flags(0, 1)
    ctx.set_config(s["X"], s["Q"])
    ctx.call("rbl_set_matvec_mode", mode)
    n = 30 * 162
    r = np.empty(3 * n, ctx.real)
    ctx.call("rbl_blob_positions", r.ctypes.data)
    F = np.random.default_rng(8).standard_normal(3 * n)
    want = _oracle_for(orc, F, r, s["a"], 1.0, True, precision)
    for v in range(ctx.L.rbl_num_sym_variants(ctx.h) if mode == 0 else ctx.L.rbl_num_matvec_variants(ctx.h)):
        ctx.call("rbl_set_sym_variant" if mode == 0 else "rbl_set_matvec_variant", v)
        check(rel_err(ctx.apply_M(F, r), want), TOL[precision])
    ctx.close()


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("n_parts", [2, 3, 8])
def test_partial_products_sum_to_the_product(orc, n_parts, precision):
    """rbl_dev_apply_M_part: the shares of the unordered-pair work a multi-GPU run hands to its
    ranks sum (all-reduce) to the full product."""
    import torch

    from rigid_body_light_b200._lib import Context

    a, eta = 0.11, 1.0
    r, F = _random_cloud(5000, True, seed=5)
    dt = _dtype(precision)
    tdt = torch.float64 if precision == "double" else torch.float32
    ctx = Context(precision)
    ctx.set_parameters(a, 0.01, 1.0, eta, np.zeros((1, 3)))
    ctx.set_flags(0, 1)
    dr = torch.from_numpy(r.reshape(-1).astype(dt)).cuda()
    dF = torch.from_numpy(F.astype(dt)).cuda()
    total = torch.zeros(3 * 5000, dtype=tdt, device="cuda")
    part = torch.empty_like(total)
    for p in range(n_parts):
        ctx.call("rbl_dev_apply_M_part", dF.data_ptr(), dr.data_ptr(), 5000, p, n_parts, part.data_ptr())
        ctx.call("rbl_sync")
        total += part
    want = _oracle_for(orc, F, r, a, eta, True, precision)
    check(rel_err(total.cpu().numpy(), want), TOL[precision])
    ctx.close()


@pytest.mark.parametrize("precision", PRECISIONS)
def test_suspension_10k_blobs_full_oracle(orc, precision):
    """64 spheres of shell_N_162 above the wall (10 368 blobs), every row against the oracle"""
    from Rigid import RigidBody
    from rigid_body_light_b200.shells import sphere_suspension

    s = sphere_suspension(64, 162, True)
    cb = RigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, precision=precision)
    r = cb.get_blob_positions()
    F = np.random.default_rng(2).standard_normal(r.size)
    out = cb.apply_M(F, r)
    check(rel_err(out, _oracle_for(orc, F, r, s["a"], 1.0, True, precision)), TOL[precision])


@pytest.fixture(scope="module")
def config2():
    """BASELINE.json configs[1]: 1000 spheres of shell_N_162 above a wall, 162 000 blobs"""
    from rigid_body_light_b200.shells import sphere_suspension

    s = sphere_suspension(1000, 162, True)
    rng = np.random.default_rng(2)
    return s, rng.standard_normal(3 * 162000), rng.standard_normal(3 * 162000)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_full_size_sampled_rows_and_properties(orc, config2, precision):
    """At the benchmark size the oracle checks 96 sampled target rows; the whole vector is
    checked through size-independent properties: symmetry <x, M y> = <M x, y> (M = B M B is
    symmetric), linearity, and run-to-run bit reproducibility (no atomics)."""
    from Rigid import RigidBody

    s, F1, F2 = config2
    cb = RigidBody(s["cfg"], s["X"], s["Q"], s["a"], 1.0, 0.01, wall_PC=True, precision=precision)
    r = cb.get_blob_positions()
    dt = _dtype(precision)
    F1, F2 = F1.astype(dt), F2.astype(dt)
    u1 = cb.apply_M(F1, r)
    u2 = cb.apply_M(F2, r)
    rows = np.random.default_rng(5).choice(162000, 96, replace=False)
    rows[:3] = [0, 161999, 81000]
    want = _oracle_for(orc, F1, r, s["a"], 1.0, True, precision, rows=rows)
    got = u1.reshape(-1, 3)[rows].reshape(-1)
    check(rel_err(got, want), TOL[precision])
    sym = abs(np.dot(F2.astype(np.float64), u1.astype(np.float64)) - np.dot(F1.astype(np.float64), u2.astype(np.float64)))
    scale = np.linalg.norm(F2) * np.linalg.norm(u1)
    assert sym / scale < (1e-12 if precision == "double" else 2e-6)
    u12 = cb.apply_M(F1 + dt(0.5) * F2, r)
    lin = rel_err(u12, u1.astype(np.float64) + 0.5 * u2.astype(np.float64))
    print(f"[{precision}] sampled-row error {rel_err(got, want):.3e}  symmetry {sym / scale:.3e}  linearity {lin:.3e}")
    assert lin < (1e-13 if precision == "double" else 5e-6)
    # the default (symmetric) kernel accumulates with floating-point atomics: reproducible to
    # rounding; the ordered kernel is bit-reproducible
    again = cb.apply_M(F1, r)
    check(rel_err(again, u1), (1e-14 if precision == "double" else 2e-6))


@pytest.mark.parametrize("precision", PRECISIONS)
def test_sharded_target_ranges_tile_the_full_product(orc, precision):
    """rbl_dev_apply_M on body-aligned target ranges (what each rank of a multi-GPU run
    computes) concatenates to the single-GPU result, bit for bit within a range layout."""
    import torch

    from rigid_body_light_b200._lib import Context
    from rigid_body_light_b200.shells import sphere_suspension

    s = sphere_suspension(48, 42, True)
    ref = s["cfg"] - s["cfg"].mean(axis=0)
    dt = _dtype(precision)
    tdt = torch.float64 if precision == "double" else torch.float32
    ctx = Context(precision)
    ctx.set_parameters(s["a"], 0.01, 1.0, 1.0, ref)
    ctx.set_flags(0, 1)
    ctx.set_config(s["X"], s["Q"])
    n = 48 * 42
    r = torch.empty(3 * n, dtype=tdt, device="cuda")
    ctx.call("rbl_dev_blob_positions", r.data_ptr())
    ctx.call("rbl_sync")
    F = torch.from_numpy(np.random.default_rng(3).standard_normal(3 * n).astype(dt)).cuda()
    full = torch.empty(3 * n, dtype=tdt, device="cuda")
    ctx.call("rbl_dev_apply_M", F.data_ptr(), r.data_ptr(), n, 0, n, full.data_ptr())
    parts = []
    for lo, hi in [(0, 13), (13, 14), (14, 40), (40, 48)]:  # body ranges
        t0, nt = lo * 42, (hi - lo) * 42
        o = torch.empty(3 * nt, dtype=tdt, device="cuda")
        ctx.call("rbl_dev_apply_M", F.data_ptr(), r.data_ptr(), n, t0, nt, o.data_ptr())
        parts.append(o)
    ctx.call("rbl_sync")
    got = torch.cat(parts).cpu().numpy()
    want = _oracle_for(orc, F.cpu().numpy(), r.cpu().numpy(), s["a"], 1.0, True, precision)
    check(rel_err(got, want), TOL[precision])
    check(rel_err(full.cpu().numpy(), want), TOL[precision])
    ctx.close()


@pytest.mark.parametrize("precision", PRECISIONS)
def test_blob_below_wall_raises(precision):
    """tests/test_wall.py:24-38"""
    from Rigid import RigidBody
    from rigid_body_light_b200.shells import icosphere_shell

    _, cfg = icosphere_shell(12)
    cb = RigidBody(cfg, np.array([[0.0, 0.0, 0.0]]), np.array([[1.0, 0, 0, 0]]), 1.0, 1.0, 1.0, wall_PC=True,
                   precision=precision)
    vec = np.random.default_rng(0).standard_normal(3 * 12 + 6)
    with pytest.raises(RuntimeError):
        cb.apply_M(vec[:36], cb.get_blob_positions())
    with pytest.raises(RuntimeError):
        cb.apply_saddle(vec)
    with pytest.raises(RuntimeError):
        cb.apply_PC(vec)
    # the context stays usable afterwards
    cb.set_config(np.array([[0.0, 0.0, 2.0]]), np.array([[1.0, 0, 0, 0]]))
    assert np.linalg.norm(cb.apply_M(vec[:36], cb.get_blob_positions())) > 0


def _full_size_sampled_rows(orc, bodies, shell, wall, precision, n_rows=64, symmetry=True):
    """One product at a BASELINE.json configuration's full size through the device-pointer C ABI:
    `n_rows` sampled target rows (first, last, seeded random) against the oracle, and -- with a second
    product -- the symmetry <F2, M F1> = <F1, M F2> of the whole vector."""
    import torch

    from rigid_body_light_b200._lib import Context
    from rigid_body_light_b200.shells import sphere_suspension

    s = sphere_suspension(bodies, shell, wall)
    ref = s["cfg"] - s["cfg"].mean(axis=0)
    ndt = _dtype(precision)
    tdt = torch.float64 if precision == "double" else torch.float32
    ctx = Context(precision)
    ctx.set_parameters(s["a"], 0.01, 1.0, 1.0, ref)
    ctx.set_flags(0, int(wall))
    ctx.set_config(s["X"], s["Q"])
    n = bodies * shell
    r = torch.empty(3 * n, dtype=tdt, device="cuda")
    ctx.call("rbl_dev_blob_positions", r.data_ptr())
    rng = np.random.default_rng(2)
    F1 = torch.from_numpy(rng.standard_normal(3 * n).astype(ndt)).cuda()
    u1 = torch.empty_like(F1)
    ctx.call("rbl_dev_apply_M", F1.data_ptr(), r.data_ptr(), n, 0, n, u1.data_ptr())
    ctx.call("rbl_sync")
    rows = np.random.default_rng(5).choice(n, n_rows, replace=False)
    rows[:2] = [0, n - 1]
    want = _oracle_for(orc, F1.cpu().numpy(), r.cpu().numpy(), s["a"], 1.0, wall, precision, rows=rows)
    got = u1.cpu().numpy().reshape(-1, 3)[rows].reshape(-1)
    err = rel_err(got, want)
    msg = f"[{bodies} x shell_N_{shell}, wall={wall}, {precision}] sampled-row error {err:.3e}"
    if symmetry:
        F2 = torch.from_numpy(rng.standard_normal(3 * n).astype(ndt)).cuda()
        u2 = torch.empty_like(F1)
        ctx.call("rbl_dev_apply_M", F2.data_ptr(), r.data_ptr(), n, 0, n, u2.data_ptr())
        ctx.call("rbl_sync")
        sym = abs(float(torch.dot(F2.double(), u1.double()) - torch.dot(F1.double(), u2.double())))
        sym /= float(F2.double().norm() * u1.double().norm())
        msg += f"  symmetry {sym:.3e}"
        assert sym < (1e-12 if precision == "double" else 2e-6), msg
    print(msg)
    assert err < TOL[precision], msg
    ctx.close()


@pytest.mark.parametrize("precision", PRECISIONS)
def test_config3_scale_wall_sampled_rows(orc, precision):
    """BASELINE.json configs[2] geometry: 4096 spheres of shell_N_42 above the wall = 172 032 blobs."""
    _full_size_sampled_rows(orc, 4096, 42, True, precision)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_config4_scale_free_space_sampled_rows(orc, precision):
    """BASELINE.json configs[3] geometry: 1000 spheres of shell_N_2562 = 2 562 000 blobs in free
    space (6.6e12 ordered pairs), float and double: 64 sampled rows against the oracle + symmetry of
    the whole product.  Exercises the 64-bit index paths of the tile triangle at full size."""
    _full_size_sampled_rows(orc, 1000, 2562, False, precision)


def test_config5_scale_wall_sampled_rows(orc):
    """BASELINE.json configs[4] geometry: 10 000 spheres of shell_N_642 = 6 420 000 blobs above the
    wall (4.1e13 ordered pairs, ~70 s on one B200 in float): the wall kernel with 64-bit triangle
    indices.  One product, 64 sampled rows against the oracle (slow, but part of the gate)."""
    _full_size_sampled_rows(orc, 10000, 642, True, "single", symmetry=False)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("wall", [False, True])
def test_apply_M_against_the_reference_members_golden(wall, precision):
    """The CUDA product against outputs of the REFERENCE'S OWN rotne_prager_tensor + apply_M
    (tests/golden/apply_M_ref_golden.npz, see tests/golden/make_golden.py): the committed cases in
    tests/test_gpu_matvec.py::test_apply_M_matches_golden use fixtures that equal these to 1e-15;
    here a ragged 257-blob cloud with overlapping and touching pairs, directly."""
    from rigid_body_light_b200._lib import Context

    ref = load_golden("apply_M_ref_golden")
    r, F, a, eta = ref["cloud/r"], ref["cloud/F"], float(ref["cloud/a"]), float(ref["cloud/eta"])
    ctx = Context(precision)
    ctx.set_parameters(a, 0.01, 1.0, eta, np.zeros((1, 3)))
    ctx.set_flags(0, wall)
    out = ctx.apply_M(F, r)
    if precision == "double":
        check(rel_err(out, ref[f"cloud/wall{int(wall)}/f64"]), TOL["double"])
    else:  # the float reference accumulates 771-term sums in float itself: compare at its own accuracy
        check(rel_err(out, ref[f"cloud/wall{int(wall)}/f64"]), TOL["single"])
        check(rel_err(out, ref[f"cloud/wall{int(wall)}/f32"]), TOL["single"])
    ctx.close()
