"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libgethi_ref.so, built
from /root/reference/src by oracle/Makefile) on small inputs.  Run in the build container only:

    OMP_NUM_THREADS=4 python tests/golden/make_golden.py

The reference realisation depends on the OpenMP thread count (per-thread MT19937, src/fourier.c:253);
the count used is recorded in each file.  Inputs: data/Pk_synth.dat (data/make_synthetic_pk.py), the
cosmology of param_GetHI_sample.ini, n+1 uniform frequency edges over 355-945 MHz written with %.6f.
"""
import ctypes as C
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.binding import Reference, write_nutable, write_param_file  # noqa: E402

HERE = Path(__file__).resolve().parent
SCALARS = ("n_grid", "l_box", "seed_rng", "do_smoothing", "r2_smooth", "fgrowth_0", "hubble_0", "numk", "logkmin",
           "logkmax", "idlogk", "n_scal", "nz_tab", "glob_idr", "dz_tab", "n_side", "n_nu", "irregular_nutable",
           "nu_min", "nu_max", "OmegaB", "hhub")


def run_case(ref, tmp, name, n_grid, n_side, n_nu, seed, full):
    nthreads = int(os.environ["OMP_NUM_THREADS"])
    nut = f"{tmp}/nu_{name}.txt"
    ini = f"{tmp}/{name}.ini"
    write_nutable(nut, n_nu)
    write_param_file(ini, n_grid=n_grid, n_side=n_side, nutable=nut, pk_file=str(ROOT / "data" / "Pk_synth.dat"),
                     prefix=f"{tmp}/{name}", seed=seed)
    par = ref.read_run_params(ini)
    d = ref.params_dict(par)
    out = {k: np.asarray(d[k]) for k in SCALARS}
    out["pos_obs"] = np.asarray(d["pos_obs"])
    for t in Reference.TABLES:
        out[t] = d[t]
    out["omp_threads"] = np.asarray(nthreads)
    for k in ("z_min", "z_max", "r_min", "r_max"):
        out[k] = np.asarray(ref.get(par, k))
    if full:
        n, nh = n_grid, n_grid // 2 + 1
        capd = np.zeros((n, n, nh), np.complex64)
        capv = np.zeros((n, n, nh), np.complex64)
        ref.lib.ref_set_fft_io(None, None, capd.ctypes.data_as(C.c_void_p), capv.ctypes.data_as(C.c_void_p))
        ref.lib.ref_create_d_and_vr_fields(par)
        ref.lib.ref_set_fft_io(None, None, None, None)
        rs = (n, n, 2 * nh)
        out["dens_k"], out["vpot_k"] = capd, capv
        out["dens"] = ref.grid(par, "dens", rs).copy()
        out["vpot"] = ref.grid(par, "vpot", rs).copy()
        out["rvel"] = ref.grid(par, "rvel", rs).copy()
        out["sigma2_gauss"] = np.asarray(ref.get(par, "sigma2_gauss"))
        ref.lib.ref_get_HI(par)
        out["mass"] = ref.grid(par, "dens", rs).copy()
        out["dz_rsd"] = ref.grid(par, "rvel", rs).copy()
        ref.lib.ref_mk_T_maps(par)
        out["maps"] = ref.grid(par, "maps_HI", (n_nu, 12 * n_side * n_side)).copy()
        # padding columns hold FFT garbage in the reference too: zero them so the file compresses
        for k in ("dens", "vpot", "rvel", "mass", "dz_rsd"):
            out[k][:, :, n:] = 0
        # reference look-ups sampled for the table-function tests
        rr = np.linspace(-5.0, 1.02 * d["r_arr_r2z"][-1], 4001)
        out["probe_r"] = rr
        out["probe_z_of_r"] = np.array([ref.lib.ref_z_of_r(par, r) for r in rr])
        out["probe_dgrowth"] = np.array([ref.lib.ref_dgrowth_of_r(par, r) for r in rr])
        out["probe_vgrowth"] = np.array([ref.lib.ref_vgrowth_of_r(par, r) for r in rr])
        zz = np.linspace(-0.1, 5.2, 2001)
        out["probe_z"] = zz
        out["probe_r_of_z"] = np.array([ref.lib.ref_r_of_z(par, z) for z in zz])
        lk = np.linspace(d["logkmin"] - 1.0, d["logkmax"] + 0.5, 3001)
        out["probe_lgk"] = lk
        out["probe_pk"] = np.array([ref.lib.ref_pk_linear0(par, x) for x in lk])
        out["probe_bias"] = np.array([ref.lib.ref_bias_HI(z) for z in zz])
        out["probe_frac"] = np.array([ref.lib.ref_fraction_HI(z) for z in zz])
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, "written", (HERE / f"{name}.npz").stat().st_size // 1024, "KiB")


def run_regular_case(tmp, name, n_grid, n_side, n_nu, seed, nu_min=355.0, nu_max=945.0):
    """The reference's other compile-time personality (no -D_IRREGULAR_NUTABLE, src/pixelize.c:176-178,216):
    uniform shells from nu_min / nu_max / n_nu and a C truncation that also puts (nu_min - dnu, nu_min) into
    shell 0.  Stores the map stage's inputs and output."""
    ref = Reference(regular=True)
    ini = f"{tmp}/{name}.ini"
    write_param_file(ini, n_grid=n_grid, n_side=n_side, regular=(nu_min, nu_max, n_nu),
                     pk_file=str(ROOT / "data" / "Pk_synth.dat"), prefix=f"{tmp}/{name}", seed=seed)
    par = ref.read_run_params(ini)
    d = ref.params_dict(par)
    assert d["irregular_nutable"] == 0
    out = {k: np.asarray(d[k]) for k in SCALARS}
    out["pos_obs"] = np.asarray(d["pos_obs"])
    for t in Reference.TABLES:
        if t in d:
            out[t] = d[t]
    out["omp_threads"] = np.asarray(int(os.environ["OMP_NUM_THREADS"]))
    n, nh = n_grid, n_grid // 2 + 1
    rs = (n, n, 2 * nh)
    ref.lib.ref_create_d_and_vr_fields(par)
    out["sigma2_gauss"] = np.asarray(ref.get(par, "sigma2_gauss"))
    ref.lib.ref_get_HI(par)
    out["mass"] = ref.grid(par, "dens", rs).copy()
    out["dz_rsd"] = ref.grid(par, "rvel", rs).copy()
    ref.lib.ref_mk_T_maps(par)
    out["maps"] = ref.grid(par, "maps_HI", (n_nu, 12 * n_side * n_side)).copy()
    for k in ("mass", "dz_rsd"):
        out[k][:, :, n:] = 0
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, "written", (HERE / f"{name}.npz").stat().st_size // 1024, "KiB")


def run_userdef_case(tmp, name, n_grid, n_side, n_nu, seed):
    """get_HI of the reference compiled with oracle/userdef_variant.c instead of its own user_defined.c (the file users
    are told to edit): inputs (Gaussian density, radial velocity, variance) and outputs (HI mass, Delta z_RSD)."""
    from oracle.binding import USERDEF_VARIANT
    ref = Reference(userdef=True)
    nut, ini = f"{tmp}/nu_{name}.txt", f"{tmp}/{name}.ini"
    write_nutable(nut, n_nu)
    write_param_file(ini, n_grid=n_grid, n_side=n_side, nutable=nut, pk_file=str(ROOT / "data" / "Pk_synth.dat"),
                     prefix=f"{tmp}/{name}", seed=seed)
    par = ref.read_run_params(ini)
    d = ref.params_dict(par)
    out = {k: np.asarray(d[k]) for k in SCALARS}
    out["pos_obs"] = np.asarray(d["pos_obs"])
    for t in Reference.TABLES:
        out[t] = d[t]
    out["omp_threads"] = np.asarray(int(os.environ["OMP_NUM_THREADS"]))
    n, nh = n_grid, n_grid // 2 + 1
    rs = (n, n, 2 * nh)
    ref.lib.ref_create_d_and_vr_fields(par)
    out["dens"] = ref.grid(par, "dens", rs).copy()
    out["rvel"] = ref.grid(par, "rvel", rs).copy()
    out["sigma2_gauss"] = np.asarray(ref.get(par, "sigma2_gauss"))
    ref.lib.ref_get_HI(par)
    out["mass"] = ref.grid(par, "dens", rs).copy()
    out["dz_rsd"] = ref.grid(par, "rvel", rs).copy()
    for k in ("dens", "rvel", "mass", "dz_rsd"):
        out[k][:, :, n:] = 0
    zz = np.linspace(0.0, 5.0, 501)
    out["probe_z"] = zz
    out["probe_bias"] = np.array([ref.lib.ref_bias_HI(z) for z in zz])
    out["probe_frac"] = np.array([ref.lib.ref_fraction_HI(z) for z in zz])
    out["userdef"] = np.array([USERDEF_VARIANT[k] for k in ("a", "p", "b0", "b1", "q")])
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, "written", (HERE / f"{name}.npz").stat().st_size // 1024, "KiB")


def main():
    """no argument: every fixture; `regular` / `userdef`: only that fixture (leaves the others untouched)."""
    if "OMP_NUM_THREADS" not in os.environ:
        raise SystemExit("set OMP_NUM_THREADS (the realisation depends on it)")
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    with tempfile.TemporaryDirectory() as tmp:
        if only in ("", "irregular"):
            ref = Reference()
            run_case(ref, tmp, "ref_n32", n_grid=32, n_side=16, n_nu=16, seed=1001, full=True)
            # tables only (they do not depend on n_grid except through l_box / pos_obs, src/cosmo.c:361-364)
            run_case(ref, tmp, "ref_tables_nu64", n_grid=512, n_side=256, n_nu=64, seed=1001, full=False)
            run_case(ref, tmp, "ref_tables_nu150", n_grid=1024, n_side=512, n_nu=150, seed=1001, full=False)
        if only in ("", "regular"):
            run_regular_case(tmp, "ref_n32_regular", n_grid=32, n_side=16, n_nu=20, seed=1001)
        if only in ("", "userdef"):
            run_userdef_case(tmp, "ref_n32_userdef", n_grid=32, n_side=16, n_nu=16, seed=1001)


if __name__ == "__main__":
    main()
