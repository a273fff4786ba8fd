"""Plane ranges of the map accumulation (host logic, no GPU): a partition of the box into contiguous ranges,
symmetric for a centred observer, wider at the box edges where most cells fall outside the shells."""
import numpy as np
import pytest

from crime_b200.gethi import params_from_tables
from crime_b200.slab import map_plane_ranges


@pytest.mark.parametrize("n_grid,nranks", [(64, 1), (64, 2), (128, 4), (256, 8), (1024, 8), (2048, 16)])
def test_partition(tables_nu150, n_grid, nranks):
    p = params_from_tables(tables_nu150, n_grid=n_grid, n_side=64)
    r = map_plane_ranges(p, nranks)
    assert r[0][0] == 0 and r[-1][1] == n_grid
    assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
    widths = np.array([hi - lo for lo, hi in r])
    assert np.abs(widths - widths[::-1]).max() <= 2          # observer at the centre
    if nranks >= 4:
        assert widths[0] > widths[nranks // 2]                  # edge ranks take more planes than centre ranks


def test_equal_model_cost(tables_nu150):
    """the ranges equalise 0.18 + 0.82 * (fraction of cells inside the shells' radial window)"""
    n, P = 256, 8
    p = params_from_tables(tables_nu150, n_grid=n, n_side=64)
    r = map_plane_ranges(p, P)
    t = tables_nu150
    h = 0.8660254 * p.l_box / n
    r_lo = np.interp(1420.40575177 / t["nuf_arr"][-1] - 1, t["z_arr_z2r"], t["r_arr_z2r"]) - h
    r_hi = np.interp(1420.40575177 / t["nu0_arr"][0] - 1, t["z_arr_z2r"], t["r_arr_z2r"]) + h
    c = p.l_box / n * (np.arange(n) + 0.5) - 0.5 * p.l_box
    rr = np.sqrt(c[:, None, None] ** 2 + c[None, :, None] ** 2 + c[None, None, :] ** 2)
    cost = 0.18 + 0.82 * ((rr > r_lo) & (rr < r_hi)).mean(axis=(1, 2))
    per_rank = np.array([cost[lo:hi].sum() for lo, hi in r])
    assert per_rank.max() / per_rank.mean() < 1.05
    slabs = cost.reshape(P, -1).sum(axis=1)
    assert slabs.max() / slabs.mean() > 1.2                     # what equal slabs would have cost


def test_off_centre_observer(tables_nu150):
    p = params_from_tables(tables_nu150, n_grid=128, n_side=64)
    p.pos_obs[2] = 0.25 * p.l_box
    r = map_plane_ranges(p, 4)
    assert r[0][0] == 0 and r[-1][1] == 128 and r[0][1] - r[0][0] < r[3][1] - r[3][0]
