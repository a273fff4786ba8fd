"""Point sources, host side: host/psources.c's setup_psources against the compiled reference's own
(src/psources.c:98-131), the tabulated luminosity distribution against the reference's rejection sampler
(draw_luminosity, src/psources.c:133-154), and temp_of_l."""
import ctypes as C

import numpy as np
import pytest

from oracle.binding import Reference, write_nutable, write_param_file

ROOT = __import__("pathlib").Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    from crime_b200 import host
    tmp = tmp_path_factory.mktemp("ps")
    write_nutable(tmp / "nu.txt", 10)
    write_param_file(tmp / "p.ini", n_grid=32, n_side=16, nutable=tmp / "nu.txt", pk_file=ROOT / "data" / "Pk_synth.dat",
                     prefix=tmp / "out", seed=5, do_psources=1)
    ref = Reference()
    par = ref.read_run_params(tmp / "p.ini")
    ref.lib.ref_setup_psources(par)
    return ref, par, host.psources_tables(tmp / "p.ini")


def test_redshift_tables_equal_the_reference(setup):
    ref, par, t = setup
    assert np.allclose(t["nz_arr"], ref.table(par, "nz_psources_arr"), rtol=1e-12, atol=0)
    assert np.allclose(t["max_Lpdf_arr"], ref.table(par, "max_Lpdf_arr"), rtol=1e-12, atol=0)
    assert abs(t["z_max"] - ref.get(par, "z_max")) < 1e-12
    nz = t["nz_arr"].size
    for z in (0.0, 0.3, 1.49, 1.51, 2.9):
        iz = int(z * nz / t["z_max"])
        zi = iz * t["z_max"] / nz
        mine = t["nz_arr"][iz] + (t["nz_arr"][iz + 1] - t["nz_arr"][iz]) * (z - zi) * nz / t["z_max"]
        assert abs(mine / ref.lib.ref_n_of_z_psources(par, z) - 1) < 1e-12


def test_tabulated_luminosity_distribution_matches_the_rejection_sampler(setup):
    """Inverse-CDF draws from the host's table against the reference's draw_luminosity: two-sample KS."""
    from scipy import stats
    ref, par, t = setup
    nz = t["nz_arr"].size
    nl = t["lcdf"].size // nz - 1
    cdf = t["lcdf"].reshape(nz, nl + 1)
    assert np.all(np.diff(cdf, axis=1) >= 0) and np.allclose(cdf[:, 0], 0) and np.allclose(cdf[:, -1], 1)
    rng = np.random.default_rng(3)
    for z in (0.4, 1.2, 2.5):
        iz = int(z * nz / t["z_max"])
        u = rng.random(100000)
        hi = np.searchsorted(cdf[iz], u, side="right")
        lo = hi - 1
        f = (u - cdf[iz][lo]) / np.maximum(cdf[iz][hi] - cdf[iz][lo], 1e-300)
        logl = t["logl_min"] + (lo + f) * (t["logl_max"] - t["logl_min"]) / nl
        theirs = np.log10(ref.draw_luminosity(par, z, 100000, seed=11))
        assert stats.ks_2samp(logl, theirs).pvalue > 1e-3
        assert abs(np.mean(10 ** logl) / np.mean(10 ** theirs) - 1) < 0.03


def test_temp_of_l_and_sed_table(setup):
    from crime_b200 import host
    ref, par, t = setup
    L = host.lib()
    L.temp_of_l.argtypes = [C.c_void_p] + [C.c_double] * 5
    L.temp_of_l.restype = C.c_double
    # the SED table reproduces spec_ed through temp_of_l's ratio at two frequencies (L0, r, z cancel)
    lognu = np.linspace(t["lognu_min"], t["lognu_max"], t["sed_arr"].size)
    for nu, z in ((400.0, 0.5), (900.0, 2.0), (600.0, 0.0)):
        a = ref.lib.ref_temp_of_l(par, 1.0, nu, z, 1000.0, 1e-3)
        sed = np.interp(np.log10((1 + z) * nu), lognu, t["sed_arr"])
        mine = 3.2548291E-2 * 8.35774E7 * 4 * np.pi * sed * t["hhub"] ** 2 / (1000.0 ** 2 * (1 + z)) / (1e-3 * nu * nu)
        assert abs(mine / a - 1) < 1e-6
