"""north_star: generated realisations must reproduce the input P(k) and a Gaussian one-point PDF within stated
statistical tolerances (tests/stat_checks.py).  The device generator is checked mode by mode against the oracle's
restatement of the same Philox stream (tests/test_gpu_parity.py) and its own 256^3 output goes through the same
checks on the GPU (test_statistical_acceptance_of_the_device_realisation); here the oracle's stream is tested on the
CPU over the eight seeds SURVEY 8(d) names."""
import numpy as np
import pytest

import stat_checks as sc
from crime_b200.gethi import params_from_tables


@pytest.mark.parametrize("seed", range(1001, 1009))
def test_power_spectrum_and_one_point_pdf(oracle, tables_nu64, seed):
    n = 64
    p = params_from_tables(tables_nu64, n_grid=n, n_side=16, seed=seed)
    dk_, vk_ = oracle.kgen_philox(p)
    dens, _, rvel, s2, mean = oracle.fields_from_k(p, dk_, vk_)
    mass, _ = oracle.get_HI(p, s2, dens, rvel)
    sc.assert_acceptance(sc.check_kspace(p, tables_nu64, dk_), sc.check_one_point(dens, s2, mean, n),
                         *sc.lognormal_mean(p, tables_nu64, mass, n))


def test_vectorised_pk_equals_the_oracle(oracle, tables_nu64):
    import ctypes as C
    p = params_from_tables(tables_nu64, n_grid=64, n_side=16)
    lg = np.linspace(p.logkmin - 0.5, p.logkmax + 0.3, 4001)
    ref = np.array([oracle.lib.oracle_pk_linear0(C.byref(p), x) for x in lg])
    assert np.abs(sc.pk_linear0_np(p, tables_nu64, lg) / ref - 1).max() < 1e-12
