"""CPU check of the arithmetic behind the per-cell Taylor path of accumulate_kernel
(crime_b200/csrc/gh_pixelize.cu): the per-cell second-order expansion of the two HEALPix ring coordinates,
evaluated in float32 exactly as the kernel does and accepted only outside the cell's confidence margin, must give
the oracle's RING pixel for every accepted sub-particle.  (The kernel itself is validated on the device by
gh_cuda_accumulate_audit; this pins the algebra and the margins without a GPU.)"""
import numpy as np
import pytest

from crime_b200.gethi import params_from_tables

F = np.float32
EPS_IDX, EPS_TT, EPS_CTH = 6e-6, 6e-6, 3e-6      # GH_FAST_EPS_* of gh_index_math.cuh


def lean_pixels(p, centres, off):
    """float32 emulation of the kernel's lean loop for cells at `centres` (double) and the ten offsets `off`."""
    ns = int(p.n_side)
    fns = F(ns)
    dx = p.l_box / p.n_grid
    x0, y0, z0 = centres.T
    xh, yh, zh = F(x0), F(y0), F(z0)
    xl, yl, zl = F(x0 - xh), F(y0 - yh), F(z0 - zh)
    rp2 = xh * xh + yh * yh
    rc = np.sqrt(zh * zh + rp2)
    h = F(dx) * F(0.8660254) + F(1e-3) + F(1e-6) * rc
    inv_rc = F(1) / rc
    irho2 = F(1) / rp2
    ir2 = inv_rc * inv_rc
    ir3 = inv_rc * ir2
    ir5 = ir3 * ir2
    k = F(0.63661977236758134308) * fns
    kq = k * irho2 * irho2
    c34 = F(0.75) * fns
    ttc = np.arctan2(yh, xh) * F(0.63661977236758134308)
    ttc = np.where(ttc < 0, ttc + F(4), ttc).astype(F)
    Ax, Ay = -k * yh * irho2, k * xh * irho2
    Axx, Axy = kq * xh * yh, kq * (yh * yh - xh * xh)
    A0 = Ax * xl + (Ay * yl + (fns * ttc + F(0.5) * fns))
    zi3, t3 = zh * ir3, F(3) * zh * ir5
    Bx, By, Bz = -c34 * xh * zi3, -c34 * yh * zi3, c34 * rp2 * ir3
    Bxx, Byy = F(0.5) * c34 * (t3 * xh * xh - zi3), F(0.5) * c34 * (t3 * yh * yh - zi3)
    Bzz = F(0.5) * c34 * (t3 * zh * zh - F(3) * zi3)
    Bxy, Bxz, Byz = c34 * t3 * xh * yh, c34 * (t3 * xh * zh - xh * ir3), c34 * (t3 * yh * zh - yh * ir3)
    B0 = Bx * xl + (By * yl + (Bz * zl + c34 * zh * inv_rc))
    dr, drho = h * inv_rc, h / np.sqrt(rp2)
    e_cell = F(EPS_IDX) * fns + fns * (F(0.2123) * drho * drho * drho + F(0.375) * dr * dr * dr)
    cth_lo = F(2.0 / 3.0) - F(EPS_CTH)
    # belt, far from the polar axis, and clear of the tt = 0 / 4 seam by more than the margins (kernel: `lean`)
    lean = ((np.abs(zh) * inv_rc + dr < cth_lo) & (rp2 > F(576) * F(dx * dx)) &
            ((xh < -h) | ((np.abs(yh) > F(2) * h) & (F(0.3) * drho * fns > fns * F(EPS_TT) + e_cell))))
    pix = np.full((len(x0), 10), -1, np.int64)
    ok_all = np.zeros((len(x0), 10), bool)
    for s in range(10):
        ox, oy, oz = F(off[s]), F(off[10 + s]), F(off[20 + s])
        # the kernel's nested evaluation (float32 sums; numpy has no fused multiply-add, which only makes this stricter)
        A = ox * (Axy * oy + (Axx * ox + Ax)) + (oy * (-Axx * oy + Ay) + A0)
        B = ox * (Bxz * oz + (Bxy * oy + (Bxx * ox + Bx))) + (oy * (Byz * oz + (Byy * oy + By)) + (oz * (Bzz * oz + Bz) + B0))
        a, b = A - B, A + B
        fa, fb = np.floor(a), np.floor(b)
        ra, rb = a - fa, b - fb
        ok = (ra > e_cell) & (ra < 1 - e_cell) & (rb > e_cell) & (rb < 1 - e_cell) & lean
        jp, jm = fa.astype(np.int64), fb.astype(np.int64)
        ir = ns + 1 + jp - jm
        ip = (jp + jm - ns + 1) >> 1                     # == (jp + jm - ns + kshift + 1) / 2 for either parity of ir
        assert np.array_equal(ip[lean], ((jp + jm - ns + 2 - (ir & 1)) >> 1)[lean])
        ip = np.where(ip >= 4 * ns, ip - 4 * ns, ip)
        pix[:, s] = 2 * ns * (ns - 1) + (ir - 1) * 4 * ns + ip
        ok_all[:, s] = ok
    return pix, ok_all, lean


@pytest.mark.parametrize("n_grid,n_side", [(512, 256), (1024, 512), (2048, 1024), (4096, 2048)])
def test_taylor_pixels_equal_the_oracle_where_accepted(oracle, tables_nu150, n_grid, n_side):
    p = params_from_tables(tables_nu150, n_grid=n_grid, n_side=n_side, seed=1001)
    rng = np.random.default_rng(n_grid)
    n = 20000
    dx = p.l_box / p.n_grid
    # random cell centres on the grid, in the shells' radial range
    idx = rng.integers(0, n_grid, (n * 6, 3))
    c = dx * (idx + 0.5) - 0.5 * p.l_box
    r = np.sqrt((c ** 2).sum(1))
    c = c[(r > float(tables_nu150["r_min"]) - 20) & (r < float(tables_nu150["r_max"]) + 20)][:n]
    off = oracle.subparticle_offsets(p)
    pix, ok, lean = lean_pixels(p, c, off)
    assert lean.mean() > 0.5                                         # most in-range cells are equatorial
    pts = (c[:, None, :] + np.stack([off[:10], off[10:20], off[20:]], axis=1)[None, :, :]).reshape(-1, 3)
    _, ref = oracle.points_to_shell_pixel(p, pts, None)
    ref = np.asarray(ref).reshape(-1, 10)
    assert ok.sum() > 0.9 * lean.sum() * 10                           # the margins reject only a few per cent
    acc = ok & (ref >= 0)                                             # the oracle reports -1 outside the shells
    assert acc.sum() > 0.8 * ok.sum()
    assert np.array_equal(pix[acc], ref[acc])                         # every accepted answer is the exact pixel
