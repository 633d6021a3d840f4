"""Equilibrium statistics of the Brownian step: spheres under gravity and a soft wall repulsion
must keep sampling the Gibbs-Boltzmann height distribution (tools/bd_equilibrium.py).  This is
the validation the reference cannot give (its step is unfinished and seeds from the wall clock,
c_rigid_obj.cpp:730-741,917-976): without the stochastic drift terms the spheres pile up at the
wall (mean height 2.29 instead of 2.36 for these parameters, about 7 standard errors of this test)."""
import os
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sedimented_spheres_sample_the_boltzmann_distribution():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bd_equilibrium

    out = bd_equilibrium.run(steps=2500, dt=0.05, side=16)  # 256 spheres, ~3000 independent samples
    want, sem = out["boltzmann_mean_h"], max(out["sem_mean_h"], 0.008)
    assert abs(out["mean_h"] - want) < 4 * sem + 0.01, out  # + O(dt) weak error of the scheme
    assert abs(out["mean_h"] - want) < 0.75 * abs(out["no_drift_mean_h"] - want), out  # on Boltzmann's side of the biased law (3 sem)
    assert 0.75 < out["var_h"] / out["boltzmann_var_h"] < 1.3, out
    assert out["min_h"] > 1.0  # nobody went through the wall
