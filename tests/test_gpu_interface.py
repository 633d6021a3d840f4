"""The reference's own behavioural test cases (/root/reference/tests/test_interface.py,
test_precision.py, test_wall.py), re-expressed against the drop-in: constructor flag
combinations, shapes, accepted dtypes, RuntimeError conditions (-m gpu)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _shell12():
    from rigid_body_light_b200.shells import icosphere_shell

    return icosphere_shell(12)[1]


def _random_bodies(n, wall=False, seed=0):
    rng = np.random.default_rng(seed)
    X = np.zeros((n, 3))
    k = 0
    while k < n:
        x = rng.uniform(1.0 if wall else -10.0, 10.0, 3)
        if k == 0 or np.all(np.linalg.norm(X[:k] - x, axis=1) > 2.0):
            X[k] = x
            k += 1
    Q = rng.standard_normal((n, 4))
    return X, Q / np.linalg.norm(Q, axis=1, keepdims=True)


def _solver(X, Q, wall_PC=False, block_PC=False, **kw):
    from Rigid import RigidBody

    return RigidBody(_shell12(), X, Q, a=1.0, eta=1.0, dt=1.0, wall_PC=wall_PC, block_PC=block_PC, **kw)


def test_create_and_bad_config():
    from Rigid import RigidBody

    cfg = _shell12()
    X = np.random.randn(10, 3)
    Q = np.random.randn(10, 4)
    RigidBody(cfg, X, Q, 1.0, 1.0, dt=0.01)
    RigidBody(cfg, X, Q, 1.0, 1.0, dt=0.01, wall_PC=True)
    RigidBody(cfg, X, Q, 1.0, 1.0, dt=0.01, block_PC=True)
    with pytest.raises(RuntimeError):
        RigidBody(cfg.flatten()[:-1], X, Q, 1.0, 1.0, dt=0.01)


def test_config_roundtrip_normalises_quaternions():
    X0 = np.random.rand(10, 3)
    Q0 = np.random.rand(10, 4)
    cb = _solver(X0, Q0)
    cb.set_config(X0, Q0)
    X, Q = cb.get_config()
    assert np.allclose(X, X0)
    assert np.allclose(Q, Q0 / np.linalg.norm(Q0, axis=1, keepdims=True))
    with pytest.raises(RuntimeError):
        cb.set_config(X0, Q0[:9])
    with pytest.raises(RuntimeError):
        cb.set_config(X0[:9], Q0)


def test_flat_inputs_keep_flat_shapes():
    X, Q = _random_bodies(3)
    cb = _solver(X.reshape(-1), Q.reshape(-1))
    assert cb.get_blob_positions().shape == (3 * 12 * 3,)
    assert cb.K_dot(np.ones(18)).shape == (108,)
    assert cb.KT_dot(np.ones(108)).shape == (18,)
    Xo, Qo = cb.get_config()
    assert Xo.shape == (9,) and Qo.shape == (12,)


def test_blob_positions_against_scipy():
    from scipy.spatial.transform import Rotation

    X, Q = _random_bodies(5)
    cfg = _shell12()
    cb = _solver(X, Q)
    pos = cb.get_blob_positions()
    assert pos.shape == (60, 3)
    want = np.concatenate([Rotation.from_quat(Q[i], scalar_first=True).apply(cfg) + X[i] for i in range(5)])
    assert np.allclose(pos, want, atol=1e-5)


def test_K_KT_sizes_and_errors():
    X, Q = _random_bodies(3)
    cb = _solver(X, Q)
    with pytest.raises(RuntimeError):
        cb.K_dot(np.random.randn(3, 3))
    with pytest.raises(RuntimeError):
        cb.KT_dot(np.random.randn(3, 3))
    out = cb.K_dot(np.random.randn(6, 3))
    assert out.shape == (36, 3) and np.linalg.norm(out) > 0
    out = cb.KT_dot(np.random.randn(36, 3))
    assert out.shape == (6, 3) and np.linalg.norm(out) > 0


def test_get_K_Kinv_are_sparse():
    X, Q = _random_bodies(3)
    cb = _solver(X, Q)
    assert np.sum(np.abs(cb.get_K())) > 0 and np.sum(np.abs(cb.get_Kinv())) > 0


@pytest.mark.parametrize("block_PC", [True, False])
@pytest.mark.parametrize("wall_PC", [True, False])
def test_apply_PC_flag_combinations(block_PC, wall_PC):
    X, Q = _random_bodies(3, wall_PC)
    cb = _solver(X, Q, wall_PC=wall_PC, block_PC=block_PC)
    n = 3 * 36 + 18
    out = cb.apply_PC(np.random.randn(n))
    assert out.shape == (n,) and np.linalg.norm(out) > 0 and np.all(np.isfinite(out))
    with pytest.raises(RuntimeError):
        cb.apply_PC(np.random.randn(n - 1))


def test_apply_M_errors_and_extra_blob():
    X, Q = _random_bodies(3)
    cb = _solver(X, Q)
    r = cb.get_blob_positions()
    with pytest.raises(RuntimeError):
        cb.apply_M(np.random.randn(35, 3), r)
    with pytest.raises(RuntimeError):
        cb.apply_M(np.random.randn(36, 3), r[:-1])
    with pytest.raises(RuntimeError):
        cb.apply_M(np.random.randn(107), np.random.randn(107))
    out = cb.apply_M(np.random.randn(36, 3), r)
    assert out.shape == (108,) and np.linalg.norm(out) > 0
    r2 = np.concatenate([r, np.random.randn(1, 3) + 20])
    assert cb.apply_M(np.random.randn(37, 3), r2).shape == (111,)


def test_apply_saddle_and_evolve():
    X, Q = _random_bodies(3)
    cb = _solver(X, Q)
    n = 108 + 18
    out = cb.apply_saddle(np.random.randn(n))
    assert out.shape == (n,) and np.linalg.norm(out) > 0
    with pytest.raises(RuntimeError):
        cb.apply_saddle(np.random.randn(n - 1))
    cb.evolve_rigid_bodies(np.random.randn(3, 6))
    Xn, Qn = cb.get_config()
    assert not np.allclose(X, Xn) and not np.allclose(Q, Qn)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("precision", ["single", "double"])
def test_both_input_dtypes_are_accepted(dtype, precision):
    """tests/test_precision.py: float32 and float64 inputs work whatever the build precision;
    outputs come back in the build precision."""
    X, Q = _random_bodies(3)
    cb = _solver(X.astype(dtype), Q.astype(dtype), precision=precision)
    want = np.float32 if precision == "single" else np.float64
    for block in (False, True):
        cb2 = _solver(X.astype(dtype), Q.astype(dtype), block_PC=block, precision=precision)
        out = cb2.apply_PC(np.random.randn(108 + 18).astype(dtype))
        assert out.dtype == want and np.linalg.norm(out) > 0
    assert cb.K_dot(np.random.randn(6, 3).astype(dtype)).dtype == want
    assert cb.KT_dot(np.random.randn(36, 3).astype(dtype)).dtype == want
    assert np.linalg.norm(cb.apply_M(np.random.randn(108).astype(dtype), cb.get_blob_positions())) > 0


def test_above_wall():
    cb = _solver(np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0, 0, 0]]), wall_PC=True)
    vec = np.random.randn(36 + 6)
    assert np.linalg.norm(cb.apply_PC(vec)) > 0
    assert np.linalg.norm(cb.apply_saddle(vec)) > 0
    assert np.linalg.norm(cb.apply_M(vec[:36], cb.get_blob_positions())) > 0
