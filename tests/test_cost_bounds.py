"""Equal-cost cuts of the unordered-pair tile triangle (include/rbl.h rbl_plan_cost_bounds): host
arithmetic only, so it runs without a GPU.  The shares of consecutive GPUs must tile the triangle
exactly, every CTA range must be well formed, and the weighted cost of the shares must be equal to
within one chunk."""
import ctypes

import numpy as np
import pytest

from rigid_body_light_b200 import _lib


def _bounds(n, tile, grid, part, parts, w):
    L = _lib.load()
    b = (ctypes.c_int64 * (grid + 1))()
    tot = ctypes.c_int64()
    assert L.rbl_plan_cost_bounds(n, tile, grid, part, parts, w, b, ctypes.byref(tot)) == 0
    return np.array(b[:], dtype=np.int64), tot.value


def _chunk_costs(n, tile, w):
    """cost of every 32-source chunk of the triangle in launch order (diagonal units first in each row)"""
    ns, D, ntt = (n + 255) // 256, tile // 256, (n + tile - 1) // tile
    out = []
    for I in range(ntt):
        length = ns - I * D
        nd = min(D, length)
        out.append(np.concatenate([np.full(8 * nd, w), np.ones(8 * (length - nd))]))
    return np.concatenate(out)


@pytest.mark.parametrize("n,tile,grid,parts", [(162000, 1536, 296, 1), (162000, 1536, 296, 8), (172032, 768, 296, 3),
                                               (5000, 256, 740, 2), (257, 256, 148, 1), (2562000, 1536, 296, 8)])
@pytest.mark.parametrize("w", [0.77, 1.0])
def test_cost_bounds_tile_the_triangle_with_equal_cost(n, tile, grid, parts, w):
    costs = _chunk_costs(n, tile, w)
    csum = np.concatenate([[0.0], np.cumsum(costs)])
    prev_end = 0
    share_cost = []
    for part in range(parts):
        b, tot = _bounds(n, tile, grid, part, parts, w)
        assert tot == costs.size
        assert b[0] == prev_end and np.all(np.diff(b) >= 0)
        prev_end = b[-1]
        piece = csum[b[1:]] - csum[b[:-1]]
        share_cost.append(piece.sum())
        # every CTA's cost within one chunk of the mean
        assert np.abs(piece - csum[-1] / (parts * grid)).max() <= 1.0 + 1e-9
    assert prev_end == costs.size
    assert max(share_cost) - min(share_cost) <= 2.0


def test_bad_arguments_are_refused():
    L = _lib.load()
    b = (ctypes.c_int64 * 4)()
    assert L.rbl_plan_cost_bounds(1000, 300, 3, 0, 1, 0.8, b, None) != 0  # tile not a multiple of 256
    assert L.rbl_plan_cost_bounds(1000, 256, 3, 2, 2, 0.8, b, None) != 0  # part out of range
