"""Binding of the C host layer (host/libgh_host.so): read_run_params -> cosmology tables, the reference-named
hot-path calls, FITS output.  The Python GetHI class (gethi.py) talks to libgh_cuda.so directly; this module
exposes what the C executable `host/GetHI` itself runs, for tests and for callers that start from a
parameter file."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import abi

HOST_LIB = abi.REPO_ROOT / "host" / "libgh_host.so"
HOST_EXE = abi.REPO_ROOT / "host" / "GetHI"
_TABLES = ("logkarr", "pkarr", "z_arr_z2r", "r_arr_z2r", "z_arr_r2z", "r_arr_r2z", "growth_d_arr", "growth_v_arr",
           "nu0_arr", "nuf_arr", "frac_HI_arr", "bias_HI_arr")
_SCALARS = ("n_grid", "l_box", "seed_rng", "do_smoothing", "r2_smooth", "fgrowth_0", "hubble_0", "numk", "logkmin",
            "logkmax", "idlogk", "n_scal", "nz_tab", "glob_idr", "dz_tab", "n_side", "n_nu", "irregular_nutable",
            "nu_min", "nu_max", "OmegaB", "hhub", "z_min", "z_max", "r_min", "r_max")
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not HOST_LIB.exists():
            raise RuntimeError(f"{HOST_LIB} not built (python -c 'import __graft_entry__ as g; g.build()')")
        abi.load_library()  # libgh_cuda.so first (RTLD_GLOBAL), libgh_host.so links against it
        L = C.CDLL(str(HOST_LIB))
        L.read_run_params_ex.argtypes = [C.c_char_p, C.c_int]
        L.read_run_params_ex.restype = C.c_void_p
        L.gh_param_double.argtypes = [C.c_void_p, C.c_char_p]
        L.gh_param_double.restype = C.c_double
        L.gh_param_table.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.gh_param_table.restype = C.POINTER(C.c_double)
        L.param_gethi_free.argtypes = [C.c_void_p]
        L.gh_write_healpix_map.argtypes = [C.c_void_p, C.c_long, C.c_char_p]
        L.gh_write_healpix_map.restype = C.c_int
        for n in ("pk_linear0", "r_of_z", "z_of_r", "dgrowth_of_r", "vgrowth_of_r"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_double]
            getattr(L, n).restype = C.c_double
        _lib = L
    return _lib


def read_run_params(fname, with_device: bool = False) -> dict:
    """Parse a GetHI parameter file with the C host code and return scalars + tables (cosmo_set included)."""
    L = lib()
    par = L.read_run_params_ex(str(fname).encode(), 1 if with_device else 0)
    d = {}
    for k in _SCALARS:
        d[k] = L.gh_param_double(par, k.encode())
    for k in ("n_grid", "seed_rng", "do_smoothing", "numk", "nz_tab", "n_side", "n_nu", "irregular_nutable"):
        d[k] = int(d[k])
    d["pos_obs"] = [L.gh_param_double(par, f"pos_obs{i}".encode()) for i in range(3)]
    for k in _TABLES:
        n = C.c_int()
        ptr = L.gh_param_table(par, k.encode(), C.byref(n))
        d[k] = np.ctypeslib.as_array(ptr, shape=(n.value,)).copy() if n.value else None
    L.param_gethi_free(par)
    return d


def psources_tables(fname) -> dict:
    """setup_psources (host/psources.c, reference src/psources.c:98-131) for a parameter file: the redshift tables of
    the reference plus the tabulated user functions that cross the C-ABI, as numpy arrays."""
    L = lib()
    L.gh_inspect_psources.argtypes = [C.c_void_p, C.POINTER(abi.GhCudaPsourcesParams)]
    par = L.read_run_params_ex(str(fname).encode(), 0)
    ps = abi.GhCudaPsourcesParams()
    L.gh_inspect_psources(par, C.byref(ps))
    as_np = lambda ptr, n: np.ctypeslib.as_array(ptr, shape=(n,)).copy()
    d = dict(nz_arr=as_np(ps.nz_arr, ps.nz), bias_arr=as_np(ps.bias_arr, ps.nz), lcdf=as_np(ps.lcdf, ps.nz * (ps.nl + 1)),
             sed_arr=as_np(ps.sed_arr, ps.nsed), z_max=ps.z_max, logl_min=ps.logl_min, logl_max=ps.logl_max, lognu_min=ps.lognu_min,
             lognu_max=ps.lognu_max, hhub=ps.hhub)
    n = C.c_int()
    d["max_Lpdf_arr"] = as_np(L.gh_param_table(par, b"max_Lpdf_arr", C.byref(n)), n.value)
    L.param_gethi_free(par)
    return d


def write_healpix_map(path, m: np.ndarray, nside: int) -> int:
    a = np.ascontiguousarray(m, dtype=np.float32)
    return lib().gh_write_healpix_map(a.ctypes.data_as(C.c_void_p), nside, str(path).encode())


def read_healpix_map(path) -> tuple[np.ndarray, dict]:
    """Minimal reader of the files write_maps produces (BINTABLE, one 1E column), for round-trip tests."""
    raw = Path(path).read_bytes()
    hdr = {}
    pos = 0
    blocks = []
    for _ in range(2):
        cards = {}
        while True:
            block = raw[pos:pos + 2880]
            pos += 2880
            done = False
            for i in range(0, 2880, 80):
                card = block[i:i + 80].decode()
                key = card[:8].strip()
                if key == "END":
                    done = True
                    break
                if card[8:10] == "= ":
                    val = card[10:].split("/")[0].strip().strip("'").strip()
                    cards[key] = val
            if done:
                break
        blocks.append(cards)
    hdr = blocks[1]
    n = int(hdr["NAXIS2"])
    data = np.frombuffer(raw[pos:pos + 4 * n], dtype=">f4").astype(np.float32)
    return data, hdr
