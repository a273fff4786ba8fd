"""Point sources on the device against the compiled reference run on the same Gaussian field
(oracle/_ref: get_point_sources src/grid_tools.c:24-101, mk_psources_maps src/pixelize.c:58-148).  Both draw from
different random streams (the reference's depends on its thread count), so: the Poisson means are compared exactly,
the catalogues and maps statistically with tolerances derived from the source counts."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
THIN = 2e-5   # thinning of n(z) so that a 64^3 box holds a few 1e5 sources (the full density gives ~1e10)


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    from crime_b200 import GetHI, abi, host
    from crime_b200.abi import GRID_DENS, GRID_RVEL
    from oracle.binding import Reference, write_nutable, write_param_file
    if not Reference.available():
        pytest.skip("oracle/_ref not built")
    tmp = tmp_path_factory.mktemp("psg")
    n, nside, n_nu = 64, 128, 10
    write_nutable(tmp / "nu.txt", n_nu)
    write_param_file(tmp / "p.ini", n_grid=n, n_side=nside, nutable=tmp / "nu.txt", pk_file=ROOT / "data" / "Pk_synth.dat",
                     prefix=tmp / "out", seed=9, do_psources=1)
    ref = Reference()
    par = ref.read_run_params(tmp / "p.ini")
    ngx = 2 * (n // 2 + 1)
    ref.lib.ref_create_d_and_vr_fields(par)
    dens = ref.grid(par, "dens", (n, n, ngx)).copy()                    # Gaussian density, before get_HI
    sigma2 = ref.get(par, "sigma2_gauss")
    ref.lib.ref_setup_psources(par)
    ref.lib.ref_scale_psources(par, THIN)
    ref.lib.ref_get_point_sources(par)
    ns_ref = ref.nsources(par, n)
    ref.lib.ref_get_HI(par)
    dz = ref.grid(par, "rvel", (n, n, ngx)).copy()                      # Delta z_RSD
    ref.lib.ref_mk_psources_maps(par)
    maps_ref = ref.maps_PS(par, n_nu, 12 * nside * nside)
    t = host.psources_tables(tmp / "p.ini")
    ps = abi.psources_params(t["nz_arr"] * THIN, t["bias_arr"], t["lcdf"], t["sed_arr"], z_max=t["z_max"], logl_min=t["logl_min"],
                             logl_max=t["logl_max"], lognu_min=t["lognu_min"], lognu_max=t["lognu_max"], hhub=t["hhub"])
    with GetHI(ref.params(par)) as g:
        g.upload_grid(GRID_DENS, dens)
        g.set_sigma2_gauss(sigma2)
        total = g.get_point_sources(ps)
        ns_dev, lam = g.download_point_sources()
        g.upload_grid(GRID_RVEL, dz)
        maps_dev = g.mk_psources_maps()
    # the Poisson means from the reference's own functions, on a sample of cells (src/grid_tools.c:62-77)
    rng = np.random.default_rng(1)
    cells = rng.integers(0, n, (400, 3))
    dx, obs = ref.get(par, "l_box") / n, ref.get(par, "pos_obs0")
    lam_ref = []
    for iz, iy, ix in cells:
        r = float(np.sqrt(sum(((c + 0.5) * dx - obs) ** 2 for c in (ix, iy, iz))))
        z = ref.lib.ref_z_of_r(par, r)
        nd = ref.lib.ref_n_of_z_psources(par, z)
        gfb = ref.lib.ref_dgrowth_of_r(par, r) * 1.0
        lam_ref.append(nd * dx ** 3 * np.exp(gfb * (float(dens[iz, iy, ix]) - 0.5 * gfb * sigma2)) if nd > 0 else 0.0)
    return dict(n=n, nside=nside, n_nu=n_nu, ns_ref=ns_ref, ns_dev=ns_dev, lam=lam, total=total, maps_ref=maps_ref, maps_dev=maps_dev,
                cells=cells, lam_ref=np.array(lam_ref))


def test_poisson_means_equal_the_reference_formula(runs):
    c = runs["cells"]
    mine = runs["lam"][c[:, 0], c[:, 1], c[:, 2]].astype(np.float64)
    ok = runs["lam_ref"] > 0
    assert ok.sum() > 100
    assert np.abs(mine[ok] / runs["lam_ref"][ok] - 1).max() < 1e-5      # float storage of the mean
    assert np.all(mine[~ok] == 0)


def test_source_counts_are_poisson_with_those_means(runs):
    lam, ns = runs["lam"].astype(np.float64), runs["ns_dev"].astype(np.float64)
    tot = lam.sum()
    assert runs["total"] == int(ns.sum()) and tot > 1e5
    assert abs(ns.sum() - tot) < 5 * np.sqrt(tot)
    assert abs(runs["ns_ref"].sum() - tot) < 5 * np.sqrt(tot)            # the reference drew from the same means
    assert np.all(ns[lam == 0] == 0)
    m = lam > 0.5
    pull = (ns[m] - lam[m]) / np.sqrt(lam[m])
    assert abs(pull.mean()) < 0.02 and abs(pull.var() - 1) < 0.05
    # same for the reference's catalogue: the two are samples of one distribution
    pull_ref = (runs["ns_ref"][m] - lam[m]) / np.sqrt(lam[m])
    assert abs(pull_ref.var() - 1) < 0.05
    # a cell-by-cell statistic that would expose a wrong sampler: P(n = 0) over cells with small means
    s = (lam > 0.05) & (lam < 0.5)
    if s.sum() > 2000:
        assert abs((ns[s] == 0).mean() - np.exp(-lam[s]).mean()) < 4 * np.sqrt(0.25 / s.sum())


def test_source_maps_match_the_reference_statistically(runs):
    a, b = runs["maps_dev"].astype(np.float64), runs["maps_ref"].astype(np.float64)
    assert a.shape == b.shape and np.all(a >= 0)
    # every source deposits in every shell: lit pixels are the same in all shells of one run
    assert np.array_equal(a[0] > 0, a[-1] > 0)
    # shell totals: sums over ~3e5 sources with var(L)/mean(L)^2 ~ 3 and a 1/r^2 weighting -> sub-per-cent scatter
    ta, tb = a.sum(1), b.sum(1)
    assert np.abs(ta / tb - 1).max() < 0.03, ta / tb
    # the spectral shape is deterministic given the catalogue: ratios of shell totals agree much better
    assert np.abs((ta / ta[0]) / (tb / tb[0]) - 1).max() < 5e-3
    # coarse angular distribution: 48 patches of ~4000 sources each.  A patch total is dominated by its few nearest /
    # most luminous sources (weights L / r^2 with r from 1300 to 4450 Mpc/h), so single patches of two independent
    # catalogues scatter by 10-20 %; the median over the patches is what has to agree
    from oracle.binding import Oracle
    orc = Oracle()
    ca, cb = orc.udgrade(a[3].astype(np.float32), 2).astype(np.float64), orc.udgrade(b[3].astype(np.float32), 2).astype(np.float64)
    assert np.abs(ca / cb - 1).max() < 1.0 and abs(np.median(ca / cb) - 1) < 0.05 and np.median(np.abs(ca / cb - 1)) < 0.15
    # the faint end is not dominated by single sources: the median lit pixel agrees well
    assert abs(np.median(a[3][a[3] > 0]) / np.median(b[3][b[3] > 0]) - 1) < 0.05
    # number of lit pixels: sources per pixel is Poisson in both
    la, lb = (a[0] > 0).mean(), (b[0] > 0).mean()
    assert abs(la - lb) < 0.02


def test_c_host_executable_with_point_sources(tmp_path):
    """do_psources=1 through host/GetHI: the source maps are written next to the HI maps (src/io_gh.c:122-128)."""
    import os
    import subprocess
    from crime_b200 import host
    from oracle.binding import write_nutable, write_param_file
    write_nutable(tmp_path / "nu.txt", 6)
    write_param_file(tmp_path / "p.ini", n_grid=64, n_side=32, nutable=tmp_path / "nu.txt", pk_file=ROOT / "data" / "Pk_synth.dat",
                     prefix=tmp_path / "run", seed=3, do_psources=1)
    r = subprocess.run([str(host.HOST_EXE), str(tmp_path / "p.ini")], capture_output=True, text=True, timeout=600,
                       env={**os.environ, "GH_PSOURCES_THIN": "1e-5"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Getting point sources" in r.stdout and "particles in total" in r.stdout
    n_src = int(r.stdout.split("There will be")[1].split()[0])
    assert 3e4 < n_src < 3e5
    tot = []
    for s in range(6):
        hi, _ = host.read_healpix_map(tmp_path / f"run_{s + 1:03d}.fits")
        ps, hdr = host.read_healpix_map(tmp_path / f"run_ps_{s + 1:03d}.fits")
        assert ps.size == 12 * 32 * 32 and np.all(ps >= 0) and ps.sum() > 0 and hi.sum() > 0
        tot.append(float(ps.astype(np.float64).sum()))
    assert tot[0] > tot[-1]            # steep synchrotron-like SED and 1 / nu^2: brighter at low frequency... shells run upward in nu
