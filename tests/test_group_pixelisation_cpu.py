"""CPU check of the grouped Taylor pixelisation of accumulate_kernel (crime_b200/csrc/gh_group_math.cuh): the very
header the kernel includes is compiled for the host (tests/native/group_math_host.cpp) and run on random 2x2x2 cell
blocks of every named configuration.  Every sub-particle the expansions accept (all rounding decisions clear of the
block's margin) must carry the oracle's RING pixel, in the equatorial belt and in both polar caps; the margins must
reject only a few per cent.  (The kernel itself is audited on the device by gh_cuda_accumulate_audit.)"""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from crime_b200.gethi import params_from_tables

HERE = Path(__file__).resolve().parent / "native"


@pytest.fixture(scope="module")
def lib():
    so = HERE / "libgroup_math_host.so"
    src = HERE / "group_math_host.cpp"
    hdr = HERE.parents[1] / "crime_b200" / "csrc" / "gh_group_math.cuh"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-o", str(so), str(src)], check=True)
    L = ctypes.CDLL(str(so))
    L.gh_group_emulate.restype = ctypes.c_int
    return L


def emulate(lib, p, centres, off_f, eps_scale=1.0):
    n = len(centres)
    kind = np.zeros(n, np.int32)
    pix = np.zeros(n * 80, np.int32)
    ok = np.zeros(n * 80, np.uint8)
    margin = np.zeros(n, np.float32)
    c = np.ascontiguousarray(centres, np.float64)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.gh_group_emulate(ctypes.c_double(p.l_box / p.n_grid), ptr(c), ctypes.c_long(n), ctypes.c_int(int(p.n_side)), ptr(off_f),
                         ctypes.c_float(eps_scale), ptr(kind), ptr(pix), ptr(ok), ptr(margin))
    return kind, pix.reshape(n, 8, 10), ok.reshape(n, 8, 10).astype(bool), margin


def block_points(p, centres, off):
    dx = p.l_box / p.n_grid
    cell = np.array([[(c & 1) - 0.5, ((c >> 1) & 1) - 0.5, ((c >> 2) & 1) - 0.5] for c in range(8)]) * dx
    sub = np.stack([off[:10], off[10:20], off[20:]], axis=1)                       # (10, 3)
    return centres[:, None, None, :] + cell[None, :, None, :] + sub[None, None, :, :]   # (n, 8, 10, 3)


@pytest.mark.parametrize("n_grid,n_side", [(512, 256), (1024, 512), (2048, 1024), (4096, 2048)])
def test_group_pixels_equal_the_oracle_where_accepted(lib, oracle, tables_nu150, n_grid, n_side):
    p = params_from_tables(tables_nu150, n_grid=n_grid, n_side=n_side, seed=1001)
    rng = np.random.default_rng(n_grid)
    n = 6000
    dx = p.l_box / p.n_grid
    # random block centres on the lattice of even cell corners, inside the shells' radial range
    idx = 2 * rng.integers(0, n_grid // 2, (n * 6, 3)) + 1
    c = dx * idx - 0.5 * p.l_box
    r = np.sqrt((c ** 2).sum(1))
    c = c[(r > float(tables_nu150["r_min"]) - 20) & (r < float(tables_nu150["r_max"]) + 20)][:n]
    off = oracle.subparticle_offsets(p)
    off_f = np.ascontiguousarray(off, np.float32)
    kind, pix, ok, margin = emulate(lib, p, c, off_f)
    pts = block_points(p, c, off).reshape(-1, 3)
    _, ref = oracle.points_to_shell_pixel(p, pts, None)
    ref = np.asarray(ref).reshape(-1, 8, 10)
    for k, name, least in ((1, "equatorial", 0.55), (2, "north", 0.1), (3, "south", 0.1)):
        sel = kind == k
        assert sel.mean() > least * (1.0 if k == 1 else 1.0), (name, sel.mean())
        acc = ok[sel] & (ref[sel] >= 0)                     # the oracle reports -1 outside the shells
        assert acc.sum() > 0.5 * ok[sel].sum()
        assert ok[sel].mean() > (0.93 if n_side <= 1024 else 0.88), (name, ok[sel].mean())
        bad = acc & (pix[sel] != ref[sel])
        assert not bad.any(), (name, int(bad.sum()), pix[sel][bad][:5], ref[sel][bad][:5])
    assert (kind > 0).mean() > 0.9                           # blocks left to the per-cell path: a few per cent


@pytest.mark.parametrize("n_grid,n_side", [(512, 256), (2048, 1024)])
def test_group_margins_have_headroom(lib, oracle, tables_nu150, n_grid, n_side):
    """With the float-evaluation margin scaled down to a quarter the accepted answers are still all exact."""
    p = params_from_tables(tables_nu150, n_grid=n_grid, n_side=n_side, seed=1001)
    rng = np.random.default_rng(7 + n_grid)
    dx = p.l_box / p.n_grid
    idx = 2 * rng.integers(0, n_grid // 2, (30000, 3)) + 1
    c = dx * idx - 0.5 * p.l_box
    r = np.sqrt((c ** 2).sum(1))
    c = c[(r > float(tables_nu150["r_min"])) & (r < float(tables_nu150["r_max"]))][:5000]
    off = oracle.subparticle_offsets(p)
    kind, pix, ok, _ = emulate(lib, p, c, np.ascontiguousarray(off, np.float32), eps_scale=0.25)
    _, ref = oracle.points_to_shell_pixel(p, block_points(p, c, off).reshape(-1, 3), None)
    ref = np.asarray(ref).reshape(-1, 8, 10)
    acc = ok & (ref >= 0) & (kind > 0)[:, None, None]
    assert not (acc & (pix != ref)).any()
