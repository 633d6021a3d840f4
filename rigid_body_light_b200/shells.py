"""Blob models (icosphere shells) and synthetic sphere suspensions.

The reference ships five blob models, ``structures/shell_N_{12,42,162,642,2562}.csv``
(header ``# sep,N,rg,rh``; parser in /root/reference/tests/utils.py:9-19).  They are
recursively subdivided icosahedra (pole vertex on +z, first ring at azimuth 72 deg)
projected onto a sphere of geometric radius ``rg`` chosen so the shell's
hydrodynamic radius is 1.  ``icosphere_shell`` regenerates them from scratch
(identical point SETS to the CSV's 8 printed decimals, checked in
tests/golden/make_golden.py against the reference files; blob ORDER is ours) so
benchmarks and GPU tests do not need the reference tree.  ``load_config`` still
reads the reference's CSV format for users who have the files.

``sphere_suspension`` builds the synthetic inputs SURVEY.md section 8(d) fixes for the
BASELINE.json configs: lattice spacing 2.5 with +-0.1 jitter, monolayer at z=1.5
above the wall (or a cube in free space), iid normal quaternions, a = sep/2,
seeds 0 (geometry) and 1 (quaternions).
"""
from __future__ import annotations

import math

import numpy as np

# N -> (subdivision level, sep, rg): parameters of the reference's shell models
SHELLS = {
    12: (0, 0.8328413657334993, 0.79207921),
    42: (1, 0.487106112144, 0.8912656),
    162: (2, 0.2620175539, 0.9496676216),
    642: (3, 0.13505535066599994, 0.9766578827),
    2562: (4, 0.06840995783799993, 0.9888262701),
}


def _icosahedron():
    z, rxy = 1.0 / math.sqrt(5.0), 2.0 / math.sqrt(5.0)
    v = [(0.0, 0.0, 1.0)]
    for k in range(5):
        t = 2 * math.pi * (k + 1) / 5
        v.append((rxy * math.cos(t), rxy * math.sin(t), z))
    for k in range(5):
        t = 2 * math.pi * (k + 1.5) / 5
        v.append((rxy * math.cos(t), rxy * math.sin(t), -z))
    v.append((0.0, 0.0, -1.0))
    f = []
    for k in range(5):
        a, b = 1 + k, 1 + (k + 1) % 5
        c, d = 6 + k, 6 + (k + 1) % 5
        f += [(0, a, b), (a, c, b), (b, c, d), (11, d, c)]
    return np.array(v), f


def icosphere(level: int) -> np.ndarray:
    """Unit-sphere vertices of an icosahedron subdivided ``level`` times
    (10*4**level + 2 points; edge midpoints re-projected at every level)."""
    v, f = _icosahedron()
    v = [tuple(p) for p in v]
    for _ in range(level):
        cache, nf = {}, []

        def mid(i, j):
            key = (i, j) if i < j else (j, i)
            if key not in cache:
                m = np.add(v[i], v[j])
                m /= np.linalg.norm(m)
                v.append(tuple(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v, dtype=np.float64)


def icosphere_shell(n_blobs: int):
    """(params, cfg) like the reference's ``load_config`` for shell_N_<n_blobs>."""
    level, sep, rg = SHELLS[n_blobs]
    cfg = icosphere(level) * rg
    return {"sep": sep, "N": n_blobs, "Rg": rg, "Rh": 1}, cfg


def load_config(file_name):
    """Reader for the reference's CSV blob models (tests/utils.py:9-19 format)."""
    with open(file_name, "r") as f:
        f.readline()
        p = f.readline().strip().lstrip("#").strip().split(",")
        cfg = np.loadtxt(f)
    return {"sep": float(p[0]), "N": int(p[1]), "Rg": float(p[2]), "Rh": float(p[3])}, cfg


def sphere_suspension(n_bodies: int, n_blobs: int, wall: bool, seed_geom: int = 0, seed_quat: int = 1):
    """Synthetic suspension of SURVEY.md section 8(d).  Returns dict with cfg, X, Q, a."""
    params, cfg = icosphere_shell(n_blobs)
    rng = np.random.default_rng(seed_geom)
    if wall:
        side = int(math.ceil(math.sqrt(n_bodies)))
        ij = np.stack(np.meshgrid(np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 2)[:n_bodies]
        X = np.concatenate([2.5 * ij, np.full((n_bodies, 1), 1.5)], axis=1).astype(np.float64)
        X[:, :2] += rng.uniform(-0.1, 0.1, (n_bodies, 2))
    else:
        side = int(math.ceil(n_bodies ** (1.0 / 3.0) - 1e-9))
        ijk = np.stack(np.meshgrid(*(np.arange(side),) * 3, indexing="ij"), -1).reshape(-1, 3)[:n_bodies]
        X = 2.5 * ijk.astype(np.float64) + rng.uniform(-0.1, 0.1, (n_bodies, 3))
    Q = np.random.default_rng(seed_quat).standard_normal((n_bodies, 4))
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    return {"cfg": cfg, "X": X, "Q": Q, "a": params["sep"] / 2.0, "params": params}
