"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol
include/rbl.h declares, fails loudly without a device (no CPU fallback), the Python API
mirror raises the reference's RuntimeErrors, and the regenerated shells match the
reference's blob models."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _cuda_present():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from rigid_body_light_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "rbl.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rbl_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 45
    L = ctypes.CDLL(_lib.lib_path())
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert b"sm_100a" in _lib.load().rbl_version()


def test_create_fails_loudly_without_a_gpu():
    if _cuda_present():
        pytest.skip("a GPU is present")
    from rigid_body_light_b200 import _lib, c_rigid

    with pytest.raises(_lib.RblError, match="no CPU fallback"):
        _lib.Context("single")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c_rigid.CManyBodies()


def test_host_classes_have_the_reference_surface():
    from Rigid import RigidBody, c_rigid  # noqa: F401  (the reference's import lines)

    assert c_rigid.CManyBodies.precision == "single"  # reference default build (pyproject.toml:14-15)
    assert c_rigid.host_class("double").precision == "double"
    bound = ["getConfig", "setParameters", "setBlkPC", "setWallPC", "set_K_mats", "K_x_U", "KT_x_Lam",
             "multi_body_pos", "apply_PC", "setConfig", "get_K", "get_Kinv", "evolve_X_Q", "apply_M", "precision"]
    for cls in (c_rigid.host_class("single"), c_rigid.host_class("double")):
        for name in bound:  # c_rigid_obj.cpp:1001-1026
            assert hasattr(cls, name), name
    for name in ["get_config", "set_config", "get_blob_positions", "K_dot", "KT_dot", "apply_M", "apply_PC",
                 "apply_saddle", "get_K", "get_Kinv", "evolve_rigid_bodies"]:  # Rigid.py:5-135
        assert hasattr(RigidBody, name), name


def test_bad_rigid_config_raises_before_touching_the_device():
    from Rigid import RigidBody

    cfg = np.zeros(35)  # not 3N  (tests/test_interface.py:21-23)
    with pytest.raises(RuntimeError, match="3N"):
        RigidBody(cfg, np.zeros((2, 3)), np.ones((2, 4)), 1.0, 1.0, dt=0.01)


def test_shells_reproduce_reference_models():
    from rigid_body_light_b200.shells import SHELLS, icosphere_shell

    chk = load_golden("shells_check")
    for n, (level, sep, rg) in SHELLS.items():
        p, cfg = icosphere_shell(n)
        assert cfg.shape == (n, 3)
        assert np.allclose(np.linalg.norm(cfg, axis=1), rg)
        assert np.abs(cfg.mean(axis=0)).max() < 1e-12
        d = np.linalg.norm(cfg[:, None, :] - cfg[None, :, :], axis=2) if n <= 642 else None
        if d is not None:
            np.fill_diagonal(d, np.inf)
            assert abs(d.min() - sep) < 2e-8  # nearest neighbours just touch at a = sep/2
        assert float(chk[f"maxdist_{n}"]) < 2e-8  # vs the reference CSV, measured when generated


def test_suspension_builder_matches_survey_recipe():
    from rigid_body_light_b200.shells import sphere_suspension

    s = sphere_suspension(1000, 162, True)
    assert s["X"].shape == (1000, 3) and np.all(s["X"][:, 2] == 1.5)
    assert abs(s["a"] - 0.2620175539 / 2) < 1e-12
    assert np.allclose(np.linalg.norm(s["Q"], axis=1), 1)
    s2 = sphere_suspension(1000, 162, True)
    assert np.array_equal(s["X"], s2["X"])  # seeded
    f = sphere_suspension(1000, 2562, False)
    assert f["X"][:, 2].max() > 20  # a cube, not a monolayer


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """include/rbl.h compiled as C99 (-pedantic -Werror) into a host that links librbl.so alone:
    no C++ or torch types cross the ABI.  Without a GPU the host sees the documented failure."""
    import subprocess

    from rigid_body_light_b200 import _lib

    exe = str(tmp_path / "capi_c_host")
    libdir = os.path.dirname(_lib.lib_path())
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host", "capi_c_host.c"), "-o", exe, "-L", libdir, "-lrbl",
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ("DEVICE-OK" in r.stdout) if _cuda_present() else ("NO-DEVICE" in r.stdout and "no CPU fallback" in r.stdout)
