"""ctypes view of the C-ABI in include/gh_cuda.h (libgh_cuda.so).

This module only describes types and loads the shared library; it has no compute of its own and
no fallback: if the CUDA library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

REPO_ROOT = Path(__file__).resolve().parent.parent
LIB_DIR = Path(__file__).resolve().parent / "csrc"
LIB_PATH = LIB_DIR / "libgh_cuda.so"

GH_CUDA_UNIQUE_ID_BYTES = 128
N_SUBPART = 10
NU_21 = 1420.40575177
GRID_DENS, GRID_VPOT, GRID_RVEL = 0, 1, 2
STAGE_NAMES = ("kgen", "fft", "vel", "sigma", "get_HI", "maps", "reduce", "d2h")

_dp = C.POINTER(C.c_double)


class GhCudaParams(C.Structure):
    """Mirror of `gh_cuda_params` (include/gh_cuda.h), itself the POD mirror of ParamGetHI
    (reference src/common_gh.h:138-209)."""

    _fields_ = [
        ("n_grid", C.c_int),
        ("l_box", C.c_double),
        ("pos_obs", C.c_double * 3),
        ("seed_rng", C.c_uint),
        ("do_smoothing", C.c_int),
        ("r2_smooth", C.c_double),
        ("fgrowth_0", C.c_double),
        ("hubble_0", C.c_double),
        ("numk", C.c_int),
        ("logkmin", C.c_double),
        ("logkmax", C.c_double),
        ("idlogk", C.c_double),
        ("n_scal", C.c_double),
        ("logkarr", _dp),
        ("pkarr", _dp),
        ("nz_tab", C.c_int),
        ("glob_idr", C.c_double),
        ("z_arr_r2z", _dp),
        ("r_arr_r2z", _dp),
        ("growth_d_arr", _dp),
        ("growth_v_arr", _dp),
        ("z_arr_z2r", _dp),
        ("r_arr_z2r", _dp),
        ("dz_tab", C.c_double),
        ("n_side", C.c_long),
        ("n_nu", C.c_int),
        ("irregular_nutable", C.c_int),
        ("nu0_arr", _dp),
        ("nuf_arr", _dp),
        ("nu_min", C.c_double),
        ("nu_max", C.c_double),
        ("OmegaB", C.c_double),
        ("hhub", C.c_double),
        ("frac_HI_arr", _dp),
        ("bias_HI_arr", _dp),
    ]


TABLE_FIELDS = ("logkarr", "pkarr", "z_arr_r2z", "r_arr_r2z", "growth_d_arr", "growth_v_arr",
                "z_arr_z2r", "r_arr_z2r", "nu0_arr", "nuf_arr", "frac_HI_arr", "bias_HI_arr")
SCALAR_FIELDS = tuple(n for n, _ in GhCudaParams._fields_ if n not in TABLE_FIELDS and n != "pos_obs")


def params_from_dict(d: dict) -> GhCudaParams:
    """Build a GhCudaParams from a dict of scalars and numpy tables.  The numpy arrays are kept
    alive on the returned struct (attribute `_keep`)."""
    p = GhCudaParams()
    keep = {}
    for name in SCALAR_FIELDS:
        if name in d:
            setattr(p, name, type(getattr(p, name))(d[name]))
    po = d["pos_obs"]
    for i in range(3):
        p.pos_obs[i] = float(po[i])
    for name in TABLE_FIELDS:
        arr = d.get(name)
        if arr is None:
            continue
        a = np.ascontiguousarray(arr, dtype=np.float64)
        keep[name] = a
        setattr(p, name, a.ctypes.data_as(_dp))
    p._keep = keep
    return p


def params_to_dict(p: GhCudaParams) -> dict:
    d = {name: getattr(p, name) for name in SCALAR_FIELDS}
    d["pos_obs"] = [p.pos_obs[i] for i in range(3)]
    for name, arr in getattr(p, "_keep", {}).items():
        d[name] = arr
    return d


class GhCudaPsourcesParams(C.Structure):
    """gh_cuda_psources_params of include/gh_cuda.h (point sources)."""
    _fields_ = [("nz", C.c_int), ("z_max", C.c_double), ("nz_arr", _dp), ("bias_arr", _dp), ("nl", C.c_int), ("logl_min", C.c_double),
                ("logl_max", C.c_double), ("lcdf", _dp), ("nsed", C.c_int), ("lognu_min", C.c_double), ("lognu_max", C.c_double),
                ("sed_arr", _dp), ("hhub", C.c_double)]


def psources_params(nz_arr, bias_arr, lcdf, sed_arr, *, z_max, logl_min, logl_max, lognu_min, lognu_max, hhub) -> GhCudaPsourcesParams:
    """Build the block from numpy tables (kept alive on the returned object)."""
    import numpy as np
    keep = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in dict(nz_arr=nz_arr, bias_arr=bias_arr, lcdf=lcdf, sed_arr=sed_arr).items()}
    nz = keep["nz_arr"].size
    p = GhCudaPsourcesParams()
    p.nz, p.z_max, p.nl, p.logl_min, p.logl_max = nz, z_max, keep["lcdf"].size // nz - 1, logl_min, logl_max
    p.nsed, p.lognu_min, p.lognu_max, p.hhub = keep["sed_arr"].size, lognu_min, lognu_max, hhub
    for k, a in keep.items():
        setattr(p, k, a.ctypes.data_as(_dp))
    p._keep = keep
    return p


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """Load libgh_cuda.so and declare every entry point of include/gh_cuda.h.  Raises if absent."""
    path = Path(path) if path else Path(os.environ.get("GH_CUDA_LIB", LIB_PATH))  # GH_CUDA_LIB: A/B-testing a build
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
    vp, i32, f64p, f32p = C.c_void_p, C.c_int, _dp, C.POINTER(C.c_float)
    u64 = C.c_ulonglong
    sig = {
        "gh_cuda_get_unique_id": ([vp], i32),
        "gh_cuda_create": ([C.POINTER(GhCudaParams), i32, i32, vp, i32, C.POINTER(vp)], i32),
        "gh_cuda_set_params": ([vp, C.POINTER(GhCudaParams)], i32),
        "gh_cuda_destroy": ([vp], i32),
        "gh_cuda_slab": ([vp, C.POINTER(i32), C.POINTER(i32)], i32),
        "gh_cuda_shells": ([vp, C.POINTER(i32), C.POINTER(i32)], i32),
        "gh_cuda_fft_pass_times": ([vp, f64p], i32),
        "gh_cuda_map_plane_bounds": ([C.POINTER(GhCudaParams), i32, C.POINTER(i32)], i32),
        "gh_cuda_create_d_and_vr_fields": ([vp, f64p, f64p], i32),
        "gh_cuda_get_HI": ([vp], i32),
        "gh_cuda_mk_T_maps": ([vp, vp], i32),
        "gh_cuda_mk_T_maps_begin": ([vp, vp], i32),
        "gh_cuda_wait_shells": ([vp, i32], i32),
        "gh_cuda_run": ([vp, f64p, vp], i32),
        "gh_cuda_run_async": ([vp, vp], i32),
        "gh_cuda_wait": ([vp, f64p], i32),
        "gh_cuda_host_alloc": ([C.POINTER(vp), u64], i32),
        "gh_cuda_host_free": ([vp], i32),
        "gh_cuda_generate_k": ([vp], i32),
        "gh_cuda_fft_fields": ([vp], i32),
        "gh_cuda_radial_velocity": ([vp], i32),
        "gh_cuda_sigma_dens": ([vp, f64p, f64p], i32),
        "gh_cuda_accumulate_maps": ([vp], i32),
        "gh_cuda_synchronize": ([vp], i32),
        "gh_cuda_set_delta_k": ([vp, vp, vp], i32),
        "gh_cuda_clear_delta_k": ([vp], i32),
        "gh_cuda_download_delta_k": ([vp, vp, vp], i32),
        "gh_cuda_download_grid": ([vp, i32, vp], i32),
        "gh_cuda_upload_grid": ([vp, i32, vp], i32),
        "gh_cuda_set_sigma2_gauss": ([vp, C.c_double], i32),
        "gh_cuda_grid_checksum": ([vp, i32, i32, i32, C.POINTER(u64)], i32),
        "gh_cuda_download_maps": ([vp, vp, u64, u64], i32),
        "gh_cuda_zero_maps": ([vp], i32),
        "gh_cuda_subparticle_offsets": ([vp, f64p], i32),
        "gh_cuda_points_to_shell_pixel": ([vp, vp, vp, C.c_longlong, vp, vp], i32),
        "gh_cuda_fastpath_audit": ([vp, vp, vp, C.c_longlong, C.c_double, vp], i32),
        "gh_cuda_accumulate_audit": ([vp, C.c_double, vp], i32),
        "gh_cuda_stage_times": ([vp, f64p], i32),
        "gh_cuda_get_point_sources": ([vp, C.POINTER(GhCudaPsourcesParams), C.POINTER(C.c_longlong)], i32),
        "gh_cuda_mk_psources_maps": ([vp, vp], i32),
        "gh_cuda_download_point_sources": ([vp, vp, vp], i32),
        "gh_cuda_jt_merge_maps": ([vp, i32, C.POINTER(vp), f64p, C.c_long, vp], i32),
        "gh_cuda_udgrade": ([vp, vp, C.c_long, vp, C.c_long, i32, i32], i32),
        "gh_cuda_nest_ring": ([vp, C.c_long, vp, vp, C.c_longlong, i32], i32),
        "gh_cuda_kernel_launches": ([vp], u64),
        "gh_cuda_stream": ([vp], vp),
        "gh_cuda_last_error": ([], C.c_char_p),
        "gh_cuda_version": ([], C.c_char_p),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = restype
    lib._gh_signatures = sig
    return lib


EXPORTED_SYMBOLS = (
    "gh_cuda_get_unique_id", "gh_cuda_create", "gh_cuda_set_params", "gh_cuda_destroy", "gh_cuda_slab", "gh_cuda_shells",
    "gh_cuda_map_plane_bounds", "gh_cuda_fft_pass_times",
    "gh_cuda_create_d_and_vr_fields", "gh_cuda_get_HI", "gh_cuda_mk_T_maps", "gh_cuda_mk_T_maps_begin", "gh_cuda_wait_shells", "gh_cuda_run", "gh_cuda_run_async", "gh_cuda_wait",
    "gh_cuda_host_alloc", "gh_cuda_host_free", "gh_cuda_generate_k", "gh_cuda_fft_fields",
    "gh_cuda_radial_velocity", "gh_cuda_sigma_dens", "gh_cuda_accumulate_maps", "gh_cuda_synchronize",
    "gh_cuda_set_delta_k", "gh_cuda_clear_delta_k", "gh_cuda_download_delta_k", "gh_cuda_download_grid",
    "gh_cuda_upload_grid", "gh_cuda_set_sigma2_gauss", "gh_cuda_grid_checksum", "gh_cuda_download_maps", "gh_cuda_zero_maps",
    "gh_cuda_subparticle_offsets", "gh_cuda_points_to_shell_pixel", "gh_cuda_fastpath_audit", "gh_cuda_accumulate_audit", "gh_cuda_stage_times",
    "gh_cuda_get_point_sources", "gh_cuda_mk_psources_maps", "gh_cuda_download_point_sources",
    "gh_cuda_jt_merge_maps", "gh_cuda_udgrade", "gh_cuda_nest_ring",
    "gh_cuda_kernel_launches", "gh_cuda_stream", "gh_cuda_last_error", "gh_cuda_version",
)
