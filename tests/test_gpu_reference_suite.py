"""The reference's OWN test-suite, unmodified, against the drop-in.

oracle/stage_ref_tests.sh copies /root/reference/tests/*.py, structures/shell_N_12.csv and
src/{__init__,Rigid}.py byte for byte into oracle/_ref/reference_tests/ (git-ignored like the
prebuilt reference libraries, travels to the GPU box with the snapshot; MANIFEST.sha256 pins the
bytes).  The suite (27 cases: tests/test_import.py:1-2, test_interface.py:8-211,
test_precision.py:7-44, test_wall.py:7-38) is run twice in a subprocess:

  layout "repo":      `Rigid` = this repository's package (Rigid/__init__.py -> RigidBody mirror)
  layout "reference": `Rigid` = the reference's own __init__.py + Rigid.py (src/Rigid.py:5-135,
                      unmodified) whose `from Rigid import c_rigid` binds this repository's
                      pybind11 CManyBodies -- the drop-in boundary of SURVEY.md section 8b.
"""
import hashlib
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

STAGED = os.path.join(ROOT, "oracle", "_ref", "reference_tests")
N_REFERENCE_CASES = 27


def _env(layout):
    env = dict(os.environ)
    paths = [ROOT]
    if layout == "reference":
        paths.insert(0, os.path.join(STAGED, "pkg_ref"))
    env["PYTHONPATH"] = os.pathsep.join(paths + [env.get("PYTHONPATH", "")]).rstrip(os.pathsep)
    env.pop("RIGID_PRECISION", None)  # the reference's default build is single precision
    return env


def _run(layout, extra):
    cmd = [sys.executable, "-m", "pytest", os.path.join(STAGED, "tests"), "-q", "-p", "no:cacheprovider",
           "--rootdir", STAGED, "-o", "addopts="] + extra
    return subprocess.run(cmd, cwd=STAGED, env=_env(layout), capture_output=True, text=True, timeout=900)


def _staged():
    return os.path.isfile(os.path.join(STAGED, "MANIFEST.sha256"))


@pytest.mark.skipif(not _staged(), reason="oracle/_ref/reference_tests not staged (run __graft_entry__.build() where /root/reference exists)")
def test_staged_files_are_the_reference_bytes():
    """MANIFEST matches the staged files; where the reference tree is present, the staged files are it."""
    pairs = {"tests/utils.py": "tests/utils.py", "tests/test_import.py": "tests/test_import.py",
             "tests/test_interface.py": "tests/test_interface.py", "tests/test_precision.py": "tests/test_precision.py",
             "tests/test_wall.py": "tests/test_wall.py", "structures/shell_N_12.csv": "structures/shell_N_12.csv",
             "pkg_ref/Rigid/__init__.py": "src/__init__.py", "pkg_ref/Rigid/Rigid.py": "src/Rigid.py"}
    manifest = {}
    for line in open(os.path.join(STAGED, "MANIFEST.sha256")):
        h, name = line.split()
        manifest[name.lstrip("./")] = h
    for staged, ref in pairs.items():
        data = open(os.path.join(STAGED, staged), "rb").read()
        assert hashlib.sha256(data).hexdigest() == manifest[staged], staged
        ref_path = os.path.join("/root/reference", ref)
        if os.path.isfile(ref_path):
            assert data == open(ref_path, "rb").read(), f"{staged} differs from {ref_path}"


@pytest.mark.skipif(not _staged(), reason="oracle/_ref/reference_tests not staged")
@pytest.mark.parametrize("layout", ["repo", "reference"])
def test_reference_suite_collects(layout):
    """No GPU needed: both layouts import and pytest finds the reference's 27 cases."""
    r = _run(layout, ["--collect-only"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    m = re.search(r"(\d+) tests? collected", r.stdout)
    assert m and int(m.group(1)) == N_REFERENCE_CASES, r.stdout[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["repo", "reference"])
def test_reference_suite_passes_unmodified(layout):
    assert _staged(), "oracle/_ref/reference_tests missing: it is staged by __graft_entry__.build() and must travel with the snapshot"
    r = _run(layout, [])
    tail = r.stdout[-4000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) == N_REFERENCE_CASES, tail
    assert "failed" not in r.stdout and "skipped" not in r.stdout, tail
    print(f"reference suite, layout={layout}: {m.group(0)}")
    if layout == "reference":  # prove which Rigid.py was driving the host class
        probe = subprocess.run([sys.executable, "-c", "import Rigid, Rigid.Rigid as R, Rigid.c_rigid as c; print(R.__file__); print(c.CManyBodies.__module__)"],
                               cwd=STAGED, env=_env(layout), capture_output=True, text=True, timeout=300)
        assert probe.returncode == 0, probe.stderr
        assert os.path.join("pkg_ref", "Rigid", "Rigid.py") in probe.stdout and "_c_rigid_f32" in probe.stdout, probe.stdout
