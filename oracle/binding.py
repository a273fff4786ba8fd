"""TEST INFRASTRUCTURE ONLY.  ctypes bindings for oracle/liboracle.so (the plain-C restatement) and
oracle/_ref/libgethi_ref.so (the unmodified reference compiled against shim/), plus the build helper.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from crime_b200.abi import GhCudaParams, N_SUBPART, params_from_dict

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "liboracle.so"
REF_SO = HERE / "_ref" / "libgethi_ref.so"
REF_SO_REGULAR = HERE / "_ref" / "libgethi_ref_regular.so"  # built without -D_IRREGULAR_NUTABLE
REF_SO_USERDEF = HERE / "_ref" / "libgethi_ref_userdef.so"  # user_defined.c swapped for oracle/userdef_variant.c
USERDEF_VARIANT = dict(a=0.012, p=0.3, b0=1.1, b1=0.07, q=2.1)  # the numbers in oracle/userdef_variant.c
REF_EXE = HERE / "_ref" / "GetHI"
REFERENCE_ROOT = Path(os.environ.get("CRIME_REFERENCE", "/root/reference"))


def build(verbose: bool = False) -> None:
    """make liboracle.so (+ _ref/ when the reference tree is present).  Building the checker is not
    using it."""
    cmd = ["make", "-C", str(HERE), "all", f"REF={REFERENCE_ROOT}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("oracle build failed")


class Slab(C.Structure):
    _fields_ = [("nz_here", C.c_int), ("iz0_here", C.c_int)]


_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_dp = C.POINTER(C.c_double)
_pp = C.POINTER(GhCudaParams)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """The plain-C restatement."""

    def __init__(self):
        if not ORACLE_SO.exists():
            build()
        self.lib = lib = C.CDLL(str(ORACLE_SO))
        for name in ("oracle_pk_linear0", "oracle_r_of_z", "oracle_z_of_r", "oracle_dgrowth_of_r",
                     "oracle_vgrowth_of_r"):
            getattr(lib, name).argtypes = [_pp, _d]
            getattr(lib, name).restype = _d
        for name in ("oracle_fraction_HI", "oracle_bias_HI"):
            getattr(lib, name).argtypes = [_d]
            getattr(lib, name).restype = _d
        lib.oracle_set_user_defined.argtypes = [_d, _d, _d, _d, _d]
        lib.oracle_kgen_mt19937.argtypes = [_pp, Slab, _i, _vp, _vp]
        lib.oracle_kgen_philox.argtypes = [_pp, _i, _i, _vp, _vp, _i]
        lib.oracle_philox4x32_10.argtypes = [_vp, _vp, _vp]
        lib.oracle_normalize.argtypes = [_pp, Slab, _vp, _vp]
        lib.oracle_radial_velocity.argtypes = [_pp, Slab, _vp, _vp, _vp, _vp]
        lib.oracle_sigma_partial.argtypes = [_pp, Slab, _vp, _dp, _dp]
        lib.oracle_fields_from_k.argtypes = [_pp, _vp, _vp, _vp, _dp, _dp]
        lib.oracle_get_HI.argtypes = [_pp, Slab, _d, _vp, _vp]
        lib.oracle_subparticle_offsets.argtypes = [_pp, _vp]
        lib.oracle_get_inu.argtypes = [_pp, _d, _i]
        lib.oracle_get_inu.restype = _i
        lib.oracle_shell_of_nu.argtypes = [_pp, _d, _i]
        lib.oracle_shell_of_nu.restype = _i
        lib.oracle_accumulate_maps.argtypes = [_pp, Slab, _vp, _vp, _vp]
        lib.oracle_normalize_maps.argtypes = [_pp, _vp]
        lib.oracle_shell_prefactors.argtypes = [_pp, _vp]
        lib.oracle_points_to_shell_pixel.argtypes = [_pp, _vp, _vp, C.c_longlong, _vp, _vp]
        lib.oracle_c2r_3d_inplace.argtypes = [_i, _vp]
        lib.oracle_fft_axis0.argtypes = [_i, _i, _i, _vp]
        lib.oracle_fft_axis1_c2r_axis2.argtypes = [_i, _i, _vp]
        lib.oracle_fft1d.argtypes = [_i, _vp]
        lib.oracle_vec2pix_ring.argtypes = [C.c_long, _vp]
        lib.oracle_vec2pix_ring.restype = C.c_long
        lib.oracle_pix2vec_ring.argtypes = [C.c_long, C.c_long, _vp]
        lib.oracle_mt_seed.argtypes = [_vp, C.c_uint32]
        lib.oracle_mt_u32.argtypes = [_vp]
        lib.oracle_mt_u32.restype = C.c_uint32

    # ---- small helpers -------------------------------------------------------------------
    @staticmethod
    def kshape(n: int, nky: int | None = None):
        return (n, n if nky is None else nky, n // 2 + 1)

    @staticmethod
    def rshape(n: int, nz: int | None = None):
        return (n if nz is None else nz, n, 2 * (n // 2 + 1))

    def mt_stream(self, seed: int, count: int) -> np.ndarray:
        state = np.zeros(625 + 8, dtype=np.uint32)
        self.lib.oracle_mt_seed(_ptr(state), seed)
        return np.array([self.lib.oracle_mt_u32(_ptr(state)) for _ in range(count)], dtype=np.uint32)

    def philox(self, ctr, key) -> np.ndarray:
        c = np.asarray(ctr, dtype=np.uint32)
        k = np.asarray(key, dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        self.lib.oracle_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
        return out

    def kgen_mt19937(self, p: GhCudaParams, n_threads: int):
        n = p.n_grid
        dk = np.zeros(self.kshape(n), dtype=np.complex64)
        vk = np.zeros(self.kshape(n), dtype=np.complex64)
        self.lib.oracle_kgen_mt19937(C.byref(p), Slab(n, 0), n_threads, _ptr(dk), _ptr(vk))
        return dk, vk

    def kgen_philox(self, p: GhCudaParams, ky0: int = 0, nky: int | None = None):
        n = p.n_grid
        nky = n if nky is None else nky
        dk = np.zeros(self.kshape(n, nky), dtype=np.complex64)
        vk = np.zeros(self.kshape(n, nky), dtype=np.complex64)
        self.lib.oracle_kgen_philox(C.byref(p), ky0, nky, _ptr(dk), _ptr(vk), 0)
        return dk, vk

    def c2r_3d(self, k: np.ndarray) -> np.ndarray:
        n = k.shape[0]
        buf = np.ascontiguousarray(k, dtype=np.complex64).copy()
        self.lib.oracle_c2r_3d_inplace(n, _ptr(buf))
        return buf.view(np.float32).reshape(self.rshape(n))

    def fields_from_k(self, p: GhCudaParams, dens_k: np.ndarray, vpot_k: np.ndarray):
        """-> dens, vpot, rvel (padded real grids), sigma2_gauss, mean_gauss."""
        n = p.n_grid
        a = np.ascontiguousarray(dens_k, dtype=np.complex64).copy()
        b = np.ascontiguousarray(vpot_k, dtype=np.complex64).copy()
        rvel = np.zeros(self.rshape(n), dtype=np.float32)
        s2, m = C.c_double(), C.c_double()
        self.lib.oracle_fields_from_k(C.byref(p), _ptr(a), _ptr(b), _ptr(rvel), C.byref(s2), C.byref(m))
        return (a.view(np.float32).reshape(self.rshape(n)), b.view(np.float32).reshape(self.rshape(n)), rvel,
                s2.value, m.value)

    def radial_velocity(self, p, vpot: np.ndarray, iz0: int, left: np.ndarray, right: np.ndarray) -> np.ndarray:
        nz = vpot.shape[0]
        out = np.zeros_like(vpot)
        self.lib.oracle_radial_velocity(C.byref(p), Slab(nz, iz0), _ptr(vpot), _ptr(left), _ptr(right), _ptr(out))
        return out

    def sigma_partial(self, p, dens: np.ndarray, iz0: int = 0):
        m, s2 = C.c_double(), C.c_double()
        self.lib.oracle_sigma_partial(C.byref(p), Slab(dens.shape[0], iz0), _ptr(dens), C.byref(m), C.byref(s2))
        return m.value, s2.value

    def get_HI(self, p, sigma2: float, dens: np.ndarray, rvel: np.ndarray, iz0: int = 0):
        d = np.ascontiguousarray(dens, dtype=np.float32).copy()
        v = np.ascontiguousarray(rvel, dtype=np.float32).copy()
        self.lib.oracle_get_HI(C.byref(p), Slab(d.shape[0], iz0), sigma2, _ptr(d), _ptr(v))
        return d, v

    def set_user_defined(self, a=0.008, p=0.6, b0=0.904, b1=0.135, q=1.696):
        """x_HI = a (1+z)^p, b_HI = b0 + b1 (1+z)^q (user_defined.c:27-35); no arguments: the shipped model."""
        self.lib.oracle_set_user_defined(a, p, b0, b1, q)

    def subparticle_offsets(self, p) -> np.ndarray:
        out = np.zeros(3 * N_SUBPART)
        self.lib.oracle_subparticle_offsets(C.byref(p), _ptr(out))
        return out

    def accumulate_maps(self, p, mass: np.ndarray, dz: np.ndarray, iz0: int = 0, maps: np.ndarray | None = None):
        npix = 12 * p.n_side * p.n_side
        if maps is None:
            maps = np.zeros((p.n_nu, npix), dtype=np.float32)
        self.lib.oracle_accumulate_maps(C.byref(p), Slab(mass.shape[0], iz0), _ptr(mass), _ptr(dz), _ptr(maps))
        return maps

    def normalize_maps(self, p, maps: np.ndarray) -> np.ndarray:
        m = maps.copy()
        self.lib.oracle_normalize_maps(C.byref(p), _ptr(m))
        return m

    def shell_prefactors(self, p) -> np.ndarray:
        out = np.zeros(p.n_nu)
        self.lib.oracle_shell_prefactors(C.byref(p), _ptr(out))
        return out

    def points_to_shell_pixel(self, p, pos: np.ndarray, dz: np.ndarray | None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = pos.shape[0]
        dzp = _ptr(np.ascontiguousarray(dz, dtype=np.float64)) if dz is not None else None
        sh = np.zeros(n, dtype=np.int32)
        px = np.zeros(n, dtype=np.int64)
        self.lib.oracle_points_to_shell_pixel(C.byref(p), _ptr(pos), dzp, n, _ptr(sh), _ptr(px))
        return sh, px

    def vec2pix_ring(self, nside: int, vec) -> int:
        v = np.asarray(vec, dtype=np.float64)
        return self.lib.oracle_vec2pix_ring(nside, _ptr(v))

    def pix2vec_ring(self, nside: int, ipix: int) -> np.ndarray:
        v = np.zeros(3)
        self.lib.oracle_pix2vec_ring(nside, ipix, _ptr(v))
        return v

    def nest2ring(self, nside: int, pix) -> np.ndarray:
        self.lib.oracle_nest2ring.restype = C.c_long
        self.lib.oracle_nest2ring.argtypes = [C.c_long, C.c_long]
        return np.array([self.lib.oracle_nest2ring(nside, int(i)) for i in np.atleast_1d(pix)], np.int64)

    def ring2nest(self, nside: int, pix) -> np.ndarray:
        self.lib.oracle_ring2nest.restype = C.c_long
        self.lib.oracle_ring2nest.argtypes = [C.c_long, C.c_long]
        return np.array([self.lib.oracle_ring2nest(nside, int(i)) for i in np.atleast_1d(pix)], np.int64)

    def udgrade(self, map_in: np.ndarray, nside_out: int, nest: bool = False) -> np.ndarray:
        """he_udgrade (src/healpix_extra.c:318-385) restated."""
        m = np.ascontiguousarray(map_in, dtype=np.float32)
        nside_in = int(round((m.size / 12) ** 0.5))
        out = np.zeros(12 * nside_out * nside_out, np.float32)
        self.lib.oracle_udgrade.argtypes = [_vp, C.c_long, _vp, C.c_long, C.c_int]
        self.lib.oracle_udgrade(_ptr(m), nside_in, _ptr(out), nside_out, int(nest))
        return out

    def run(self, p: GhCudaParams, dens_k=None, vpot_k=None):
        """Whole hot path on one slab; Philox stream unless a k-space field is supplied."""
        if dens_k is None:
            dens_k, vpot_k = self.kgen_philox(p)
        dens, vpot, rvel, s2, mean = self.fields_from_k(p, dens_k, vpot_k)
        mass, dz = self.get_HI(p, s2, dens, rvel)
        maps = self.normalize_maps(p, self.accumulate_maps(p, mass, dz))
        return dict(dens=dens, vpot=vpot, rvel=rvel, sigma2=s2, mean=mean, mass=mass, dz=dz, maps=maps)


class Reference:
    """The unmodified reference (oracle/_ref/libgethi_ref.so) behind ref_harness.c."""

    TABLES = ("logkarr", "pkarr", "z_arr_z2r", "r_arr_z2r", "z_arr_r2z", "r_arr_r2z", "growth_d_arr",
              "growth_v_arr", "nu0_arr", "nuf_arr")

    # -- point sources (do_psources = 1) --
    def nsources(self, par, n_grid: int) -> np.ndarray:
        """nsources as get_point_sources left it, [N][N][N] (the reference indexes it with the padded pitch)."""
        ngx = 2 * (n_grid // 2 + 1)
        a = np.ctypeslib.as_array(self.lib.ref_nsources(par), shape=(n_grid, n_grid, ngx))
        return a[:, :, :n_grid].copy()

    def maps_PS(self, par, n_nu: int, npix: int) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.ref_maps_PS(par), shape=(n_nu, npix)).copy()

    def draw_luminosity(self, par, z: float, n: int, seed: int = 7) -> np.ndarray:
        out = np.zeros(n)
        self.lib.ref_draw_luminosity(par, z, seed, n, out.ctypes.data_as(_vp))
        return out

    @staticmethod
    def available(regular: bool = False) -> bool:
        return (REF_SO_REGULAR if regular else REF_SO).exists()

    def __init__(self, regular: bool = False, userdef: bool = False):
        """regular=True: the build without -D_IRREGULAR_NUTABLE (param keys nu_min / nu_max / n_nu);
        userdef=True: the build with oracle/userdef_variant.c in place of the reference's user_defined.c."""
        so = REF_SO_USERDEF if userdef else (REF_SO_REGULAR if regular else REF_SO)
        self.regular = regular
        if not so.exists():
            raise RuntimeError(f"{so} missing (needs the reference tree to build; see oracle/Makefile)")
        self.lib = lib = C.CDLL(str(so))
        lib.ref_read_run_params.argtypes = [C.c_char_p]
        lib.ref_read_run_params.restype = _vp
        for n in ("ref_create_d_and_vr_fields", "ref_get_HI", "ref_mk_T_maps", "ref_write_maps", "ref_free"):
            getattr(lib, n).argtypes = [_vp]
            getattr(lib, n).restype = None
        for n in ("ref_pk_linear0", "ref_z_of_r", "ref_r_of_z", "ref_dgrowth_of_r", "ref_vgrowth_of_r"):
            getattr(lib, n).argtypes = [_vp, _d]
            getattr(lib, n).restype = _d
        for n in ("ref_fraction_HI", "ref_bias_HI"):
            getattr(lib, n).argtypes = [_d]
            getattr(lib, n).restype = _d
        lib.ref_get_double.argtypes = [_vp, C.c_char_p]
        lib.ref_get_double.restype = _d
        lib.ref_set_double.argtypes = [_vp, C.c_char_p, _d]
        lib.ref_get_table.argtypes = [_vp, C.c_char_p, C.POINTER(_i)]
        lib.ref_get_table.restype = _dp
        lib.ref_grid.argtypes = [_vp, C.c_char_p]
        lib.ref_grid.restype = C.POINTER(C.c_float)
        lib.ref_set_fft_io.argtypes = [_vp, _vp, _vp, _vp]
        for n in ("ref_setup_psources", "ref_get_point_sources", "ref_mk_psources_maps"):
            getattr(lib, n).argtypes = [_vp]
            getattr(lib, n).restype = None
        lib.ref_scale_psources.argtypes = [_vp, _d]
        lib.ref_n_of_z_psources.argtypes = [_vp, _d]
        lib.ref_n_of_z_psources.restype = _d
        lib.ref_temp_of_l.argtypes = [_vp, _d, _d, _d, _d, _d]
        lib.ref_temp_of_l.restype = _d
        lib.ref_draw_luminosity.argtypes = [_vp, _d, C.c_uint, _i, _vp]
        lib.ref_draw_luminosity.restype = _d
        lib.ref_nsources.argtypes = [_vp]
        lib.ref_nsources.restype = C.POINTER(C.c_int)
        lib.ref_maps_PS.argtypes = [_vp]
        lib.ref_maps_PS.restype = C.POINTER(C.c_float)

    def read_run_params(self, fname: str):
        """The reference's -D_DEBUG cosmo_set dumps test_cosmo.dat into the current directory: run it from a
        scratch directory so that the working tree stays clean (paths in the parameter file must be absolute)."""
        import tempfile
        fname = str(Path(fname).resolve())
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as scratch:
            os.chdir(scratch)
            try:
                return self.lib.ref_read_run_params(fname.encode())
            finally:
                os.chdir(cwd)

    def get(self, par, name: str) -> float:
        return self.lib.ref_get_double(par, name.encode())

    def table(self, par, name: str) -> np.ndarray:
        n = _i()
        ptr = self.lib.ref_get_table(par, name.encode(), C.byref(n))
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()

    def grid(self, par, name: str, shape) -> np.ndarray:
        ptr = self.lib.ref_grid(par, name.encode())
        return np.ctypeslib.as_array(ptr, shape=tuple(shape))

    def params_dict(self, par) -> dict:
        g = lambda k: self.get(par, k)
        d = dict(
            n_grid=int(g("n_grid")), l_box=g("l_box"), pos_obs=[g("pos_obs0"), g("pos_obs1"), g("pos_obs2")],
            seed_rng=int(g("seed_rng")), do_smoothing=int(g("do_smoothing")), r2_smooth=g("r2_smooth"),
            fgrowth_0=g("fgrowth_0"), hubble_0=g("hubble_0"), numk=int(g("numk")), logkmin=g("logkmin"),
            logkmax=g("logkmax"), idlogk=g("idlogk"), n_scal=g("n_scal"), nz_tab=5001, glob_idr=g("glob_idr"),
            dz_tab=0.001, n_side=int(g("n_side")), n_nu=int(g("n_nu")), irregular_nutable=0 if self.regular else 1,
            nu_min=g("nu_min"), nu_max=g("nu_max"), OmegaB=g("OmegaB"), hhub=g("hhub"))
        for t in self.TABLES:
            if self.regular and t in ("nu0_arr", "nuf_arr"):
                continue
            d[t] = self.table(par, t)
        return d

    def params(self, par) -> GhCudaParams:
        return params_from_dict(self.params_dict(par))


def write_param_file(path, *, n_grid, n_side, pk_file, prefix, nutable=None, regular=None, seed=1001, r_smooth=2.0,
                     omega_M=0.3, omega_L=0.7, omega_B=0.049, h=0.67, w=-1.0, ns=0.96, sigma_8=0.8, do_psources=0):
    """A GetHI param file with the reference's keys (param_GetHI_sample.ini).  nutable: file of shell edges (the
    -D_IRREGULAR_NUTABLE personality); regular=(nu_min, nu_max, n_nu): the other one (src/io_gh.c:234-241)."""
    freq = (f"frequencies_filename= {nutable}\n" if regular is None else
            f"nu_min= {regular[0]}\nnu_max= {regular[1]}\nn_nu= {regular[2]}\n")
    Path(path).write_text(
        f"prefix_out= {prefix}\npk_filename= {pk_file}\nomega_M= {omega_M}\nomega_L= {omega_L}\n"
        f"omega_B= {omega_B}\nh= {h}\nw= {w}\nns= {ns}\nsigma_8= {sigma_8}\nr_smooth= {r_smooth}\n"
        f"{freq}n_side= {n_side}\nn_grid= {n_grid}\nseed= {seed}\n"
        f"do_psources= {do_psources}\n")


def write_nutable(path, n_nu: int, nu_min: float = 355.0, nu_max: float = 945.0):
    edges = np.linspace(nu_min, nu_max, n_nu + 1)
    Path(path).write_text("".join(f"{e:.6f}\n" for e in edges))
