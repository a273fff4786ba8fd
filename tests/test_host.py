"""Host C layer (host/): parameter file parsing, cosmology tables against the golden vectors produced by
the reference's own cosmo_set, FITS writer round trip.  CPU only."""
import subprocess

import numpy as np
import pytest

from crime_b200 import host
from oracle.binding import write_nutable, write_param_file
from conftest import ROOT


@pytest.fixture(scope="module")
def parsed(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("host")
    write_nutable(tmp / "nu.txt", 64)
    write_param_file(tmp / "p.ini", n_grid=512, n_side=256, nutable=tmp / "nu.txt", pk_file=ROOT / "data" / "Pk_synth.dat",
                     prefix=tmp / "out")
    return host.read_run_params(tmp / "p.ini")


def test_scalars_match_reference(parsed, tables_nu64):
    g = tables_nu64
    for k in ("n_grid", "n_side", "n_nu", "numk", "seed_rng", "do_smoothing", "irregular_nutable"):
        assert parsed[k] == int(g[k]), k
    assert parsed["r2_smooth"] == 4.0  # r_smooth squared in place (src/io_gh.c:255-260)
    for k in ("l_box", "fgrowth_0", "hubble_0", "glob_idr", "logkmin", "logkmax", "idlogk", "z_min", "z_max", "r_min", "r_max"):
        assert parsed[k] == pytest.approx(float(g[k]), rel=2e-7), k
    assert parsed["pos_obs"] == pytest.approx(list(g["pos_obs"]), rel=2e-7)


def test_tables_match_reference(parsed, tables_nu64):
    """The reference integrates with GSL qng/qagil at relative tolerances 1e-6 (distances) and 1e-4 (growth,
    sigma_8); the host code integrates to ~1e-10, so agreement with the golden tables (made with a tight
    integrator standing in for GSL) is at the 1e-7 level."""
    g = tables_nu64
    for k in ("z_arr_z2r", "r_arr_z2r", "z_arr_r2z", "r_arr_r2z", "growth_d_arr", "growth_v_arr", "logkarr", "nu0_arr", "nuf_arr"):
        a, b = parsed[k], g[k]
        assert a.shape == b.shape, k
        assert np.abs(a - b).max() <= 3e-7 * np.abs(b).max(), k
    assert np.abs(parsed["pkarr"] / g["pkarr"] - 1).max() < 2e-6  # sigma_8 normalisation integral


def test_regular_nutable_keys(tmp_path):
    (tmp_path / "p.ini").write_text(
        f"prefix_out= {tmp_path}/o\npk_filename= {ROOT}/data/Pk_synth.dat\nnu_min= 400.\nnu_max= 800.\nn_nu= 20\n"
        "n_grid= 64\nn_side= 16\nseed= 7\nr_smooth= -1\nbogus_key= 3\n# a comment\n\ndo_psources= 0\n")
    d = host.read_run_params(tmp_path / "p.ini")
    assert d["irregular_nutable"] == 0 and d["n_nu"] == 20 and d["nu_min"] == 400. and d["nu_max"] == 800.
    assert d["do_smoothing"] == 0 and d["seed_rng"] == 7
    assert d["nu0_arr"] is None


def test_fits_round_trip(tmp_path):
    nside = 8
    m = np.random.default_rng(0).standard_normal(12 * nside * nside).astype(np.float32)
    f = tmp_path / "m_001.fits"
    assert host.write_healpix_map(f, m, nside) == 0
    assert f.stat().st_size % 2880 == 0
    back, hdr = host.read_healpix_map(f)
    assert np.array_equal(back, m)
    assert hdr["PIXTYPE"] == "HEALPIX" and hdr["ORDERING"] == "RING" and int(hdr["NSIDE"]) == nside
    assert hdr["TFORM1"] == "1E" and hdr["TUNIT1"] == "mK" and hdr["COORDSYS"] == "G" and hdr["TTYPE1"] == "T"
    assert host.write_healpix_map(f, m, nside) == 1  # never overwrites silently


def test_cli_usage_and_missing_gpu(tmp_path):
    r = subprocess.run([str(host.HOST_EXE)], capture_output=True, text=True)
    assert r.returncode == 0 and "Usage: ./GetHI file_name" in r.stderr


def _tiny_param_file(tmp_path):
    write_nutable(tmp_path / "nu.txt", 4)
    write_param_file(tmp_path / "p.ini", n_grid=32, n_side=8, nutable=tmp_path / "nu.txt", pk_file=ROOT / "data" / "Pk_synth.dat",
                     prefix=tmp_path / "out")
    return tmp_path / "p.ini"


def _no_gpu_here():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu_here(), reason="checks the failure path of a box without a GPU")
def test_fork_launcher_fails_loudly_without_a_gpu(tmp_path):
    """GH_NGPUS=2 ./GetHI on a box with no GPU: rank 0 cannot make an NCCL id (or a context); the other rank must see
    end-of-file on its pipe instead of waiting for ever, and the launcher must return non-zero promptly."""
    import os
    ini = _tiny_param_file(tmp_path)
    r = subprocess.run([str(host.HOST_EXE), str(ini)], capture_output=True, text=True, timeout=60,
                       env={**os.environ, "GH_NGPUS": "2"})
    assert r.returncode != 0
    assert "Fatal" in r.stderr


@pytest.mark.skipif(not _no_gpu_here(), reason="checks the failure path of a box without a GPU")
def test_launcher_environment_path_without_a_gpu(tmp_path):
    """RANK / WORLD_SIZE / GH_UNIQUE_ID_FILE (torchrun-style launch): a non-zero rank that never gets an id file gives
    up with a message; without GH_UNIQUE_ID_FILE the program says what is missing."""
    import os
    ini = _tiny_param_file(tmp_path)
    env = {k: v for k, v in os.environ.items() if k not in ("GH_NGPUS", "GH_UNIQUE_ID_FILE")}
    r = subprocess.run([str(host.HOST_EXE), str(ini)], capture_output=True, text=True, timeout=60,
                       env={**env, "GH_RANK": "1", "GH_NRANKS": "2"})
    assert r.returncode != 0 and "GH_UNIQUE_ID_FILE" in r.stderr
    r = subprocess.run([str(host.HOST_EXE), str(ini)], capture_output=True, text=True, timeout=60,
                       env={**env, "GH_RANK": "0", "GH_NRANKS": "2", "GH_UNIQUE_ID_FILE": str(tmp_path / "id")})
    assert r.returncode != 0 and "Fatal" in r.stderr


def test_user_defined_hooks_are_tabulated_for_the_device(parsed):
    """host/user_defined.c (the reference's src/user_defined.c:27-35): cosmo_set samples both hooks at z_arr_r2z and
    hotpath.c hands the samples across the C-ABI."""
    z = parsed["z_arr_r2z"]
    assert np.allclose(parsed["frac_HI_arr"], 0.008 * (1 + z) ** 0.6, rtol=1e-14)
    assert np.allclose(parsed["bias_HI_arr"], 0.904 + 0.135 * (1 + z) ** 1.696, rtol=1e-14)
