"""CPU transcription of the radix plan, the decimation-in-frequency passes, the in-register butterflies and the
digit reversal of crime_b200/csrc/gh_fft.cu (one line, double precision), checked against numpy for every line
length the library instantiates -- in particular 1024, 2048 and 4096, whose kernels only run at sizes no unit
test on a GPU reaches.  What this pins: n_eights / n_fours / rad_at / rad_prod, dft<4> / dft<8> output order,
the per-pass twiddle index i*(N/M), and dif_pos_to_freq.  (Thread mapping, shared-memory addressing and the
fused transposes are the same code for every length and are covered on the device at N <= 512.)"""
import numpy as np
import pytest


def ilog2(n):
    return 0 if n <= 1 else 1 + ilog2(n >> 1)


def n_eights(n):
    return ilog2(n) // 3 - 1 if ilog2(n) % 3 == 1 else ilog2(n) // 3


def n_fours(n):
    return 2 if ilog2(n) % 3 == 1 else (1 if ilog2(n) % 3 == 2 else 0)


def n_steps(n):
    return n_eights(n) + n_fours(n)


def rad_at(n, s):
    return 8 if s < n_eights(n) else 4


def rad_prod(n, s):
    return 1 if s <= 0 else rad_prod(n, s - 1) * rad_at(n, s - 1)


def mul_pi(a):
    return complex(-a.imag, a.real)


def dft4(u0, u1, u2, u3):
    a, b, c, d = u0 + u2, u0 - u2, u1 + u3, mul_pi(u1 - u3)
    return a + c, b + d, a - c, b - d


def dft(u):
    if len(u) == 4:
        return list(dft4(*u))
    e = dft4(u[0], u[2], u[4], u[6])
    o = dft4(u[1], u[3], u[5], u[7])
    h = 0.70710678118654752440
    o1 = complex((o[1].real - o[1].imag) * h, (o[1].real + o[1].imag) * h)
    o3 = complex((-o[3].real - o[3].imag) * h, (o[3].real - o[3].imag) * h)
    oo = (o[0], o1, mul_pi(o[2]), o3)
    out = [0j] * 8
    for k in range(4):
        out[k], out[k + 4] = e[k] + oo[k], e[k] - oo[k]
    return out


def dif_pos_to_freq(n, pos):
    f = 0
    for s in range(n_steps(n)):
        sub, r, mult = n // rad_prod(n, s + 1), rad_at(n, s), rad_prod(n, s)
        f += ((pos // sub) % r) * mult
    return f


def kernel_line_fft(x):
    n = len(x)
    x = np.array(x, dtype=np.complex128)
    tw = np.exp(2j * np.pi * np.arange(n) / n)
    ns = n_steps(n)
    for s in range(ns):
        r, m = rad_at(n, s), n // rad_prod(n, s)
        sub = m // r
        for b in range(n // m):
            for i in range(sub):
                p0 = b * m + i
                u = dft([x[p0 + k * sub] for k in range(r)])
                if sub > 1:
                    w1 = tw[i * (n // m)]
                    u = [u[q] * w1 ** q for q in range(r)]
                for q in range(r):
                    x[p0 + q * sub] = u[q]
    out = np.empty_like(x)
    r_last = rad_at(n, ns - 1)
    for b in range(n // r_last):
        f0 = dif_pos_to_freq(n, b * r_last)
        for q in range(r_last):
            # the last pass scatters u[q] to f0 + q*(N/R)
            assert dif_pos_to_freq(n, b * r_last + q) == f0 + q * (n // r_last)
            out[f0 + q * (n // r_last)] = x[b * r_last + q]
    return out


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_radix_plan_and_digit_reversal(n):
    assert rad_prod(n, n_steps(n)) == n and n_eights(n) >= 0
    assert sorted(dif_pos_to_freq(n, p) for p in range(n)) == list(range(n))     # a permutation
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    got = kernel_line_fft(x)
    ref = np.fft.ifft(x) * n                                                     # exponent sign +, unnormalised
    assert np.abs(got - ref).max() < 1e-9 * np.abs(ref).max()


def fft_cfg(n):
    """FftCfg of gh_fft.cu"""
    w = 16 if n <= 512 else (8 if n <= 2048 else 4)
    nt_s0 = min(512, max(64, w * n // 16))
    nt_s = 256 if (n <= 1024 and nt_s0 > 256) else nt_s0
    wr = 16 if n <= 1024 else (8 if n <= 2048 else 4)
    nt_r = min(512, max(64, wr * (n // 2) // 16))
    return w, nt_s, wr, nt_r


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_tile_shapes_fit_the_sm(n):
    """shared-memory tiles of both kernels fit 227 KB with at least one CTA per SM, thread counts are whole warps,
    and a tile row is at least one 32-byte sector"""
    w, nt_s, wr, nt_r = fft_cfg(n)
    assert n * w * 8 <= 227 * 1024 and (n // 2) * wr * 8 <= 227 * 1024
    assert nt_s % 32 == 0 and nt_r % 32 == 0 and 64 <= nt_s <= 512 and 64 <= nt_r <= 512
    assert w * 8 >= 32
    # the half-length transform of the x pass needs a plan of its own
    assert n_steps(n // 2) >= 1 and rad_prod(n // 2, n_steps(n // 2)) == n // 2
