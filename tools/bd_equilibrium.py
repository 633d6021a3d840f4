"""Statistical validation of the Brownian step (rbl_bd_step): independent-potential spheres above
the wall must sample the Gibbs-Boltzmann distribution of their height.

side^2 spheres (shell_N_12, hydrodynamic radius 1) on a lattice of spacing 6 above the wall,
each under gravity m g and a soft wall repulsion  U(h) = m g h + eps exp(-(h - R) / b).  The
potential is one-body, so whatever the hydrodynamic coupling the equilibrium marginal of every
height is  P(h) ~ exp(-U(h) / kBT).  A scheme without the stochastic drift kBT d(mu)/dh would
sample P(h) / mu_perp(h) instead, which piles the spheres onto the wall (mu_perp -> 0 there):
the check has power.  It probes the translational drift (RFD of M + midpoint); orientation
dependent drift of non-spherical bodies is not exercised by spheres.

Prints one JSON line: measured mean / variance of h against the Boltzmann values."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def boltzmann(mg, eps, b, R, kBT):
    """grid, normalised density, mean and variance of P(h) ~ exp(-(m g h + eps exp(-(h-R)/b)) / kBT)"""
    g = np.linspace(R - 0.2, R + 14.0, 400001)
    w = np.exp(-(mg * g + eps * np.exp(-(g - R) / b)) / kBT)
    Z = np.trapezoid(w, g)
    m1 = np.trapezoid(w * g, g) / Z
    m2 = np.trapezoid(w * g * g, g) / Z
    return g, w / Z, m1, m2 - m1 * m1


def run(steps=4000, burn=0, dt=0.05, kBT=1.0, mg=2.0, eps=32.0, b=0.25, R=1.0, side=8, precision="double", seed=5):
    from Rigid import RigidBody
    from rigid_body_light_b200.shells import icosphere_shell

    params, cfg = icosphere_shell(12)
    a = params["sep"] / 2.0
    nb = side * side
    rng = np.random.default_rng(seed)
    ij = np.stack(np.meshgrid(np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 2)
    # start IN equilibrium (inverse-CDF samples of the Boltzmann heights): a correct scheme stays
    # there from step 0, a wrong drift walks away from it
    g, pdf, m1, var = boltzmann(mg, eps, b, R, kBT)
    cdf = np.cumsum(pdf) * (g[1] - g[0])
    h0 = np.interp(rng.uniform(0.002, 0.998, nb), cdf, g)
    X = np.concatenate([6.0 * ij, h0[:, None]], axis=1).astype(np.float64)
    Q = rng.standard_normal((nb, 4))
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    cb = RigidBody(cfg, X, Q, a, 1.0, dt, wall_PC=True, block_PC=True, precision=precision)
    n3 = 3 * nb * 12
    hs = []
    t0 = time.perf_counter()
    for k in range(steps):
        Xc, _ = cb.get_config()
        h = Xc[:, 2].astype(np.float64)
        F = np.zeros((nb, 6))
        F[:, 2] = -mg + (eps / b) * np.exp(-(h - R) / b)
        noise = tuple(rng.standard_normal(n3) for _ in range(3))
        cb.bd_step(F.reshape(-1), kBT=kBT, noise=noise, tol=1e-6, restart=40, max_iter=80, lanczos_tol=1e-4,
                   lanczos_max_iter=36)
        if k >= burn:
            hs.append(h)
    wall_s = time.perf_counter() - t0
    hs = np.array(hs)
    per_body = hs.mean(axis=0)
    # what a scheme WITHOUT the stochastic drift would sample: P(h) / mu_perp(h), with the
    # single-sphere wall mobility mu_perp/mu_0 ~ 1 - 9/(8h) + 1/(2h^3) (the far-field form the
    # wall correction encodes, c_rigid_obj.cpp:104), clipped near contact
    mu = np.clip(1 - 9 / (8 * g) + 1 / (2 * g ** 3), 0.05, None)
    wb = pdf / mu
    biased_mean = float(np.trapezoid(wb * g, g) / np.trapezoid(wb, g))
    return {"steps": steps, "burn": burn, "dt": dt, "kBT": kBT, "mg": mg, "eps": eps, "b": b, "bodies": nb,
            "precision": precision, "mean_h": float(hs.mean()), "boltzmann_mean_h": float(m1),
            "var_h": float(hs.var()), "boltzmann_var_h": float(var), "no_drift_mean_h": biased_mean,
            "sem_mean_h": float(per_body.std(ddof=1) / np.sqrt(nb)), "min_h": float(hs.min()),
            "seconds": wall_s, "ms_per_step": 1e3 * wall_s / steps}


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    side = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    print(json.dumps(run(steps=steps, side=side)), flush=True)
