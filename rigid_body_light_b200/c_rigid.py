"""Selector for the compiled host class, importable as ``Rigid.c_rigid``.

The reference compiles ONE precision per build (``real`` = float unless
-DDOUBLEPRECISION=1, /root/reference/src/eigen_defines.h:5-37, pyproject.toml:14-15)
and exposes it as ``CManyBodies.precision``.  Here both builds ship; the module-level
default follows the reference (single) unless ``RIGID_PRECISION=double`` is set, and
``host_class("double")`` gives the other one explicitly.

There is no CPU fallback: if the extension modules (or librbl.so behind them) are
missing, importing this module raises ImportError telling the user to run
``python -c "import __graft_entry__ as g; g.build()"``.
"""
import os

try:
    from . import _c_rigid_f32, _c_rigid_f64
except ImportError as exc:  # pragma: no cover - build problem, not a code path
    raise ImportError(
        "rigid_body_light_b200: the CUDA host class is not built (run "
        "`make -C rigid_body_light_b200/csrc` or __graft_entry__.build()); "
        "there is no CPU fallback. Original error: %s" % (exc,)
    ) from exc

_BY_NAME = {"single": _c_rigid_f32, "double": _c_rigid_f64,
            "float32": _c_rigid_f32, "float64": _c_rigid_f64}


def host_class(precision=None):
    """``CManyBodies`` class for ``precision`` in {"single","double"} (None = default)."""
    if precision is None:
        precision = os.environ.get("RIGID_PRECISION", "single")
    try:
        return _BY_NAME[str(precision).lower()].CManyBodies
    except KeyError:
        raise RuntimeError(f"unknown precision {precision!r}; use 'single' or 'double'") from None


CManyBodies = host_class()
precision = CManyBodies.precision
