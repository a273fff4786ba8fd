"""CPU prototype of the second-order (per-cell Taylor) equatorial pixelisation for accumulate_kernel: checks the
algebra and measures the truncation error that the confidence margins must cover.  Not part of the product."""
import numpy as np

rng = np.random.default_rng(1)


def coeffs(x, y, z, ns):
    """Taylor coefficients about the cell centre of A = ns*tt + ns/2 and B = 0.75*ns*z/r (float64)."""
    rho2 = x * x + y * y
    r2 = rho2 + z * z
    ir = 1 / np.sqrt(r2)
    ir3, ir5 = ir ** 3, ir ** 5
    k = 2 / np.pi * ns
    tt = np.arctan2(y, x) * 2 / np.pi
    tt = np.where(tt < 0, tt + 4, tt)
    A = dict(c=ns * tt + 0.5 * ns, x=-k * y / rho2, y=k * x / rho2, xx=k * x * y / rho2 ** 2, yy=-k * x * y / rho2 ** 2,
             xy=k * (y * y - x * x) / rho2 ** 2)
    c = 0.75 * ns
    B = dict(c=c * z * ir, x=-c * x * z * ir3, y=-c * y * z * ir3, z=c * rho2 * ir3,
             xx=0.5 * c * (-z * ir3 + 3 * z * x * x * ir5), yy=0.5 * c * (-z * ir3 + 3 * z * y * y * ir5),
             zz=0.5 * c * (-3 * z * ir3 + 3 * z ** 3 * ir5), xy=c * 3 * z * x * y * ir5,
             xz=c * (-x * ir3 + 3 * x * z * z * ir5), yz=c * (-y * ir3 + 3 * y * z * z * ir5))
    return A, B


def exact(x, y, z, ns):
    tt = np.arctan2(y, x) * 2 / np.pi
    tt = np.where(tt < 0, tt + 4, tt)
    return ns * tt + 0.5 * ns, 0.75 * ns * z / np.sqrt(x * x + y * y + z * z)


n = 400000
ns = 256
for dx, rmin, rmax in ((17.45, 1300.0, 4500.0), (8.7, 1300.0, 4500.0)):
    r = rng.uniform(rmin, rmax, n)
    cth = rng.uniform(-0.68, 0.68, n)
    ph = rng.uniform(0.05, 2 * np.pi - 0.05, n)
    st = np.sqrt(1 - cth * cth)
    x, y, z = r * st * np.cos(ph), r * st * np.sin(ph), r * cth
    o = dx * (rng.random((3, n)) - 0.5)
    A, B = coeffs(x, y, z, ns)
    At = A["c"] + A["x"] * o[0] + A["y"] * o[1] + A["xx"] * o[0] ** 2 + A["yy"] * o[1] ** 2 + A["xy"] * o[0] * o[1]
    Bt = (B["c"] + B["x"] * o[0] + B["y"] * o[1] + B["z"] * o[2] + B["xx"] * o[0] ** 2 + B["yy"] * o[1] ** 2 + B["zz"] * o[2] ** 2
          + B["xy"] * o[0] * o[1] + B["xz"] * o[0] * o[2] + B["yz"] * o[1] * o[2])
    Ae, Be = exact(x + o[0], y + o[1], z + o[2], ns)
    d = np.sqrt((o ** 2).sum(0))
    rho = np.sqrt(x * x + y * y)
    eA, eB = np.abs(At - Ae), np.abs(Bt - Be)
    print(f"dx={dx}: max |A err| {eA.max():.3e} px, as multiple of (2/pi) ns (d/rho)^3/3: {(eA / (2 / np.pi * ns * (d / rho) ** 3 / 3)).max():.3f}")
    print(f"          max |B err| {eB.max():.3e} px, as multiple of 0.75 ns (d/r)^3: {(eB / (0.75 * ns * (d / r) ** 3)).max():.3f}")


def f32(v):
    return np.asarray(v, dtype=np.float32)


print("\nfloat32 evaluation (coefficients and sums in float32) vs float64 exact:")
for ns, dx in ((256, 17.45), (512, 8.7), (1024, 4.35), (2048, 2.17)):
    r = rng.uniform(1300.0, 4500.0, n)
    cth = rng.uniform(-0.66, 0.66, n)
    ph = rng.uniform(0.05, 2 * np.pi - 0.05, n)
    st = np.sqrt(1 - cth * cth)
    xd, yd, zd = r * st * np.cos(ph), r * st * np.sin(ph), r * cth
    x, y, z = f32(xd), f32(yd), f32(zd)                      # hi parts
    lo = np.stack([xd - x, yd - y, zd - z])                  # lo parts (double remainder)
    o = dx * (rng.random((3, n)) - 0.5)
    of = f32(o)
    # coefficients in float32 from the hi parts
    with np.errstate(all="ignore"):
        rho2 = x * x + y * y
        r2 = rho2 + z * z
        ir = f32(1) / np.sqrt(r2)
        ir2 = ir * ir
        ir3 = ir * ir2
        ir5 = ir3 * ir2
        irho2 = f32(1) / rho2
        k = f32(2 / np.pi * ns)
        tt = np.arctan2(y, x) * f32(2 / np.pi)
        tt = np.where(tt < 0, tt + f32(4), tt)
        A0 = f32(ns) * tt + f32(0.5 * ns)
        Ax, Ay = -k * y * irho2, k * x * irho2
        Axx = k * x * y * irho2 * irho2
        Axy = k * (y * y - x * x) * irho2 * irho2
        c = f32(0.75 * ns)
        B0 = c * z * ir
        Bx, By, Bz = -c * x * z * ir3, -c * y * z * ir3, c * rho2 * ir3
        t3 = f32(3) * z * ir5
        Bxx = f32(0.5) * c * (t3 * x * x - z * ir3)
        Byy = f32(0.5) * c * (t3 * y * y - z * ir3)
        Bzz = f32(0.5) * c * (t3 * z * z - f32(3) * z * ir3)
        Bxy = c * t3 * x * y
        Bxz = c * (t3 * x * z - x * ir3)
        Byz = c * (t3 * y * z - y * ir3)
        # lo parts folded into the constants through the gradient
        A0 = A0 + Ax * f32(lo[0]) + Ay * f32(lo[1])
        B0 = B0 + Bx * f32(lo[0]) + By * f32(lo[1]) + Bz * f32(lo[2])
        ox, oy, oz = of
        At = A0 + Ax * ox + Ay * oy + Axx * (ox * ox - oy * oy) + Axy * (ox * oy)
        Bt = B0 + Bx * ox + By * oy + Bz * oz + Bxx * (ox * ox) + Byy * (oy * oy) + Bzz * (oz * oz) + Bxy * (ox * oy) + Bxz * (ox * oz) + Byz * (oy * oz)
    Ae, Be = exact(xd + of[0].astype(np.float64), yd + of[1].astype(np.float64), zd + of[2].astype(np.float64), ns)
    print(f"ns={ns:5d} dx={dx:5.2f}: max|A err| {np.abs(At - Ae).max():.3e} px  max|B err| {np.abs(Bt - Be).max():.3e} px   "
          f"(margin 6e-6*ns = {6e-6 * ns:.3e})")
