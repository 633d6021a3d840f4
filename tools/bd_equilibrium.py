"""Statistical validation of the Brownian step (rbl_bd_step): independent-potential spheres above
the wall must sample the Gibbs-Boltzmann distribution of their height.

64 spheres (shell_N_12, hydrodynamic radius 1) on an 8 x 8 lattice of spacing 6 above the wall,
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


def run(steps=3000, burn=300, dt=0.02, kBT=1.0, mg=2.0, eps=8.0, b=0.25, R=1.0, side=8, precision="double", seed=5):
    from Rigid import RigidBody
    from rigid_body_light_b200.shells import icosphere_shell

    params, cfg = icosphere_shell(12)
    a = params["sep"] / 2.0
    nb = side * side
    rng = np.random.default_rng(seed)
    ij = np.stack(np.meshgrid(np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 2)
    X = np.concatenate([6.0 * ij, np.full((nb, 1), 2.0)], axis=1).astype(np.float64)
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
    # Boltzmann moments by quadrature
    g = np.linspace(R - 0.2, R + 12.0, 200001)
    w = np.exp(-(mg * g + eps * np.exp(-(g - R) / b)) / kBT)
    Z = np.trapezoid(w, g)
    m1 = np.trapezoid(w * g, g) / Z
    m2 = np.trapezoid(w * g * g, g) / Z
    # the biased (no-drift) distribution for comparison: P / mu_perp with the Brenner-like fit of the
    # reference's single-blob wall self-mobility is not needed; report the plain numbers
    per_body = hs.mean(axis=0)
    return {"steps": steps, "burn": burn, "dt": dt, "kBT": kBT, "mg": mg, "eps": eps, "b": b, "bodies": nb,
            "precision": precision, "mean_h": float(hs.mean()), "boltzmann_mean_h": float(m1),
            "var_h": float(hs.var()), "boltzmann_var_h": float(m2 - m1 * m1),
            "sem_mean_h": float(per_body.std(ddof=1) / np.sqrt(nb)), "min_h": float(hs.min()),
            "seconds": wall_s, "ms_per_step": 1e3 * wall_s / steps}


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    print(json.dumps(run(steps=steps, burn=max(100, steps // 10))), flush=True)
