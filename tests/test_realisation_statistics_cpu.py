"""north_star: generated realisations must reproduce the input P(k) and a Gaussian one-point PDF within stated
statistical tolerances.  The device generator is checked mode by mode against the oracle's restatement of the same
Philox stream (tests/test_gpu_parity.py::test_kgen_matches_oracle_philox), so the statistics of that stream are
tested here on the CPU, over the eight seeds SURVEY 8(d) names: shell-averaged power within 4 sigma of
P(k) exp(-r_s^2 k^2) / dk^3 in every k bin (mode-count errors; kx = 0 and Nyquist planes included -- every stored
mode is an independent draw of the same variance), phases uniform, and the real-space field Gaussian with the
variance the path reports (Kolmogorov-Smirnov on a sparse sub-lattice)."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

from crime_b200.gethi import params_from_tables


@pytest.mark.parametrize("seed", range(1001, 1009))
def test_power_spectrum_and_one_point_pdf(oracle, tables_nu64, seed):
    n = 64
    p = params_from_tables(tables_nu64, n_grid=n, n_side=16, seed=seed)
    dk_, _ = oracle.kgen_philox(p)
    dk = 2 * np.pi / p.l_box
    idx = np.fft.fftfreq(n, 1.0 / n)
    kz, ky, kx = np.meshgrid(idx, idx, np.arange(n // 2 + 1), indexing="ij")
    k2 = (kx ** 2 + ky ** 2 + kz ** 2) * dk * dk
    sel = k2 > 0
    lg = 0.5 * np.log10(k2[sel])
    var = np.array([oracle.lib.oracle_pk_linear0(C.byref(p), x) for x in lg]) / dk ** 3 * np.exp(-p.r2_smooth * k2[sel])
    ratio = np.abs(dk_[sel]) ** 2 / var                  # exponential(1) per mode for a Rayleigh modulus
    kmod = np.sqrt(k2[sel])
    edges = np.linspace(0, kmod.max() * 1.0001, 25)
    which = np.digitize(kmod, edges) - 1
    worst = 0.0
    for b in range(24):
        r = ratio[which == b]
        if r.size < 30:
            continue
        worst = max(worst, abs(r.mean() - 1) * np.sqrt(r.size))   # in sigma: var of exp(1) is 1
    assert worst < 4.0
    # phases uniform in [0, 2 pi)
    ph = np.angle(dk_[sel]) % (2 * np.pi)
    assert stats.kstest(ph / (2 * np.pi), "uniform").pvalue > 1e-3
    # real-space one-point PDF: Gaussian with the variance the path reports
    dens, _, _, s2, mean = oracle.fields_from_k(p, dk_, np.zeros_like(dk_))
    sub = dens[::4, ::4, :n:4].ravel().astype(np.float64)
    assert abs(mean) < 1e-6
    assert stats.kstest((sub - mean) / np.sqrt(s2), "norm").pvalue > 1e-3
    assert abs(stats.skew(sub)) < 0.15 and abs(stats.kurtosis(sub)) < 0.3
