import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure; never imported by the package)."""
    from oracle import oracle

    oracle.build()
    return oracle


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


CASE_NAMES = ["case_overlap_wall", "case_overlap_free", "case_touch_wall", "case_touch_free", "case_near_wall"]


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


# tolerances of BASELINE.json's north_star: <= 1e-12 relative in double, <= 1e-5 in float
TOL = {"double": 1e-12, "single": 1e-5}


# ---- parity bookkeeping ------------------------------------------------------------------
# Every parity assertion goes through check(): it asserts err < tol AND records the observed
# error with its call site, so that the tolerances written in the tests can be held within 10x of
# what the code actually does ("tighten the parity asserts").  The record is printed
# at the end of the session and, on the GPU box, written to gpurun_out/parity_observed.jsonl.
_OBSERVED = []


def check(err, tol, what=None):
    import inspect
    import json

    fr = inspect.stack()[1]
    site = f"{os.path.basename(fr.filename)}:{fr.lineno}"
    test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0].split("::")[-1]
    err, tol = float(err), float(tol)
    _OBSERVED.append({"site": site, "test": test, "what": what, "err": err, "tol": tol})
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_observed.jsonl"), "a") as fh:
            fh.write(json.dumps(_OBSERVED[-1]) + "\n")
    assert err < tol, f"{what or site}: observed {err:.3e} >= tolerance {tol:.1e} ({test})"


def pytest_terminal_summary(terminalreporter):
    if not _OBSERVED:
        return
    worst = {}
    for o in _OBSERVED:
        k = (o["site"], o["tol"])
        worst[k] = max(worst.get(k, 0.0), o["err"])
    terminalreporter.write_line("parity: worst observed error per assertion site (tolerance)")
    for (site, tol), e in sorted(worst.items()):
        terminalreporter.write_line(f"  {site:32s} {e:10.3e}  ({tol:.1e})")
