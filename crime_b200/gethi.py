"""Host-side mirror of the reference's GetHI hot-path interface over the C-ABI (include/gh_cuda.h).

The reference drives the path with five calls on one state struct (src/main_gh.c:24-80):
``init_fftw``, ``create_d_and_vr_fields``, ``get_HI``, ``mk_T_maps``, ``end_fftw``.  `GetHI` keeps those
names and their order; the work itself happens in libgh_cuda.so (hand-written sm_100a kernels).  There is
no CPU implementation behind this class: without the CUDA library or a GPU every call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import abi
from .abi import GhCudaParams, GRID_DENS, GRID_RVEL, GRID_VPOT, STAGE_NAMES

_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = abi.load_library()
    return _LIB


class GetHIError(RuntimeError):
    """Raised where the reference would print `Node %d, Fatal: ...` and exit(1) (src/common_gh.c:104-120)."""


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def params_from_tables(tables, *, n_grid: int, n_side: int, seed: int = 1001, r_smooth: float | None = None) -> GhCudaParams:
    """ParamGetHI for a given grid from the cosmology / P(k) / frequency tables `cosmo_set` produces
    (src/cosmo.c:341-413).  Only l_box and pos_obs depend on the grid (src/cosmo.c:361-364)."""
    t = {k: tables[k] for k in tables.keys()} if not isinstance(tables, dict) else dict(tables)
    d = {k: (t[k].item() if np.ndim(t[k]) == 0 else t[k]) for k in t}
    r_max = float(d["r_max"])
    l_box = 2 * r_max * (1 + 2.0 / n_grid)
    d.update(n_grid=n_grid, n_side=n_side, seed_rng=seed, l_box=l_box, pos_obs=[0.5 * l_box] * 3)
    if r_smooth is not None:
        d["do_smoothing"] = int(r_smooth > 0)
        d["r2_smooth"] = r_smooth * r_smooth if r_smooth > 0 else r_smooth
    return abi.params_from_dict(d)


class GetHI:
    """One rank's view of the run: owns z planes [iz0_here, iz0_here+nz_here) and, after mk_T_maps, the
    shells [shell0_here, shell0_here+n_shells_here)."""

    def __init__(self, params: GhCudaParams, rank: int = 0, nranks: int = 1, unique_id: bytes | None = None,
                 device: int = 0):
        self.lib = lib()
        self.params = params
        self.rank, self.nranks = rank, nranks
        self._ctx = C.c_void_p()
        self.sigma2_gauss = -1.0
        self.mean_gauss = 0.0
        self.maps_HI = None
        self._pinned = None
        self.init_fftw(unique_id, device)

    # -- the five reference entry points ---------------------------------------------------------
    def init_fftw(self, unique_id: bytes | None = None, device: int = 0) -> None:
        """src/fourier.c:101 (+ mpi_init, allocate_maps): slab bounds and device allocations."""
        uid = C.create_string_buffer(unique_id, abi.GH_CUDA_UNIQUE_ID_BYTES) if unique_id is not None else None
        self._check(self.lib.gh_cuda_create(C.byref(self.params), self.rank, self.nranks, uid, device,
                                            C.byref(self._ctx)))
        nz, iz0, ns, s0 = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.gh_cuda_slab(self._ctx, C.byref(nz), C.byref(iz0)))
        self._check(self.lib.gh_cuda_shells(self._ctx, C.byref(ns), C.byref(s0)))
        self.nz_here, self.iz0_here = nz.value, iz0.value
        self.n_shells_here, self.shell0_here = ns.value, s0.value
        self.n_grid = self.params.n_grid
        self.npix = 12 * self.params.n_side ** 2

    def set_params(self, params: GhCudaParams) -> None:
        """Hand a fresh parameter block (same sizes) to the device: what read_run_params does before a run."""
        self._check(self.lib.gh_cuda_set_params(self._ctx, C.byref(params)))
        self.params = params

    def create_d_and_vr_fields(self) -> float:
        """src/fourier.c:375-438.  Returns sigma2_gauss."""
        s2, m = C.c_double(), C.c_double()
        self._check(self.lib.gh_cuda_create_d_and_vr_fields(self._ctx, C.byref(s2), C.byref(m)))
        self.sigma2_gauss, self.mean_gauss = s2.value, m.value
        return s2.value

    def get_HI(self) -> None:
        """src/grid_tools.c:103-153."""
        self._check(self.lib.gh_cuda_get_HI(self._ctx))

    def mk_T_maps(self, to_host: bool = True) -> np.ndarray | None:
        """src/pixelize.c:150-286.  Returns this rank's shells, [n_shells_here][npix] float32 (host)."""
        if not to_host:
            self._check(self.lib.gh_cuda_mk_T_maps(self._ctx, None))
            return None
        buf = self._host_maps()
        self._check(self.lib.gh_cuda_mk_T_maps(self._ctx, _ptr(buf)))
        self.maps_HI = buf
        return buf

    def mk_T_maps_begin(self, slot: int = 0) -> np.ndarray:
        """Non-blocking mk_T_maps: the download of this rank's shells is queued in chunks of whole shells;
        wait_shells(n) returns when the first n are complete in the returned buffer."""
        buf = self._host_maps(slot)
        self._check(self.lib.gh_cuda_mk_T_maps_begin(self._ctx, _ptr(buf)))
        self.maps_HI = buf
        return buf

    def wait_shells(self, n_shells: int = -1) -> None:
        self._check(self.lib.gh_cuda_wait_shells(self._ctx, n_shells))

    def end_fftw(self) -> None:
        """src/fourier.c:201 (+ the grid part of param_gethi_free, src/io_gh.c:298-324)."""
        if self._ctx:
            if self._pinned:
                self.lib.gh_cuda_wait(self._ctx, None)
                for p, _ in self._pinned.values():
                    self.lib.gh_cuda_host_free(p)
                self._pinned = None
                self.maps_HI = None
            self.lib.gh_cuda_destroy(self._ctx)
            self._ctx = C.c_void_p()

    close = end_fftw

    def run(self, to_host: bool = True):
        """main_gh.c:52-62 back to back."""
        s2 = C.c_double()
        buf = self._host_maps() if to_host else None
        self._check(self.lib.gh_cuda_run(self._ctx, C.byref(s2), _ptr(buf)))
        self.sigma2_gauss = s2.value
        self.maps_HI = buf
        return buf

    def run_async(self, slot: int | None = 0) -> np.ndarray | None:
        """Enqueue a whole realisation without waiting; the maps land in pinned host buffer `slot` (0 or 1; None: they
        stay on the device).  Call wait() before reading them."""
        if slot is None:
            self._check(self.lib.gh_cuda_run_async(self._ctx, None))
            return None
        buf = self._host_maps(slot)
        self._check(self.lib.gh_cuda_run_async(self._ctx, _ptr(buf)))
        return buf

    def wait(self) -> float:
        s2 = C.c_double()
        self._check(self.lib.gh_cuda_wait(self._ctx, C.byref(s2)))
        self.sigma2_gauss = s2.value
        return s2.value

    # -- finer stages ------------------------------------------------------------------------------
    def generate_k(self):
        self._check(self.lib.gh_cuda_generate_k(self._ctx))

    def fft_fields(self):
        self._check(self.lib.gh_cuda_fft_fields(self._ctx))

    def radial_velocity(self):
        self._check(self.lib.gh_cuda_radial_velocity(self._ctx))

    def sigma_dens(self):
        s2, m = C.c_double(), C.c_double()
        self._check(self.lib.gh_cuda_sigma_dens(self._ctx, C.byref(s2), C.byref(m)))
        self.sigma2_gauss, self.mean_gauss = s2.value, m.value
        return s2.value, m.value

    def accumulate_maps(self):
        self._check(self.lib.gh_cuda_accumulate_maps(self._ctx))

    def zero_maps(self):
        self._check(self.lib.gh_cuda_zero_maps(self._ctx))

    def synchronize(self):
        self._check(self.lib.gh_cuda_synchronize(self._ctx))

    # -- injection / read-back ---------------------------------------------------------------------
    def kshape(self):
        n = self.n_grid
        return (n, n, n // 2 + 1)

    def slab_shape(self):
        n = self.n_grid
        return (self.nz_here, n, 2 * (n // 2 + 1))

    def set_delta_k(self, dens_k: np.ndarray, vpot_k: np.ndarray):
        a = np.ascontiguousarray(dens_k, dtype=np.complex64)
        b = np.ascontiguousarray(vpot_k, dtype=np.complex64)
        assert a.shape == self.kshape() and b.shape == self.kshape()
        self._check(self.lib.gh_cuda_set_delta_k(self._ctx, _ptr(a), _ptr(b)))

    def clear_delta_k(self):
        self._check(self.lib.gh_cuda_clear_delta_k(self._ctx))

    def download_delta_k(self):
        a = np.zeros(self.kshape(), np.complex64)
        b = np.zeros(self.kshape(), np.complex64)
        self._check(self.lib.gh_cuda_download_delta_k(self._ctx, _ptr(a), _ptr(b)))
        return a, b

    def download_grid(self, which: int) -> np.ndarray:
        out = np.zeros(self.slab_shape(), np.float32)
        self._check(self.lib.gh_cuda_download_grid(self._ctx, which, _ptr(out)))
        return out

    def upload_grid(self, which: int, slab: np.ndarray):
        a = np.ascontiguousarray(slab, dtype=np.float32)
        assert a.shape == self.slab_shape()
        self._check(self.lib.gh_cuda_upload_grid(self._ctx, which, _ptr(a)))

    def set_sigma2_gauss(self, s2: float):
        self._check(self.lib.gh_cuda_set_sigma2_gauss(self._ctx, float(s2)))
        self.sigma2_gauss = float(s2)

    def grid_checksum(self, which: int, z0_local: int = 0, n_planes: int | None = None) -> int:
        """Position-weighted checksum of the real cells of planes [z0_local, z0_local+n_planes) of this slab."""
        out = C.c_ulonglong()
        n_planes = self.nz_here - z0_local if n_planes is None else n_planes
        self._check(self.lib.gh_cuda_grid_checksum(self._ctx, which, z0_local, n_planes, C.byref(out)))
        return int(out.value)

    def download_maps(self) -> np.ndarray:
        """Full per-rank stack as it sits on the device before any cross-rank reduction."""
        out = np.zeros((self.params.n_nu, self.npix), np.float32)
        self._check(self.lib.gh_cuda_download_maps(self._ctx, _ptr(out), 0, out.size))
        return out

    def subparticle_offsets(self) -> np.ndarray:
        out = np.zeros(3 * abi.N_SUBPART)
        self._check(self.lib.gh_cuda_subparticle_offsets(self._ctx, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def points_to_shell_pixel(self, pos: np.ndarray, dz: np.ndarray | None = None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = pos.shape[0]
        dzc = np.ascontiguousarray(dz, dtype=np.float64) if dz is not None else None
        sh = np.zeros(n, np.int32)
        px = np.zeros(n, np.int64)
        self._check(self.lib.gh_cuda_points_to_shell_pixel(self._ctx, _ptr(pos), _ptr(dzc), n, _ptr(sh), _ptr(px)))
        return sh, px

    def fastpath_audit(self, pos: np.ndarray, dz: np.ndarray | None = None, eps_scale: float = 1.0) -> dict:
        """mk_T_maps' fp32 fast path vs its exact path on the given points."""
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        dzc = np.ascontiguousarray(dz, dtype=np.float64) if dz is not None else None
        cnt = np.zeros(4, np.uint64)
        self._check(self.lib.gh_cuda_fastpath_audit(self._ctx, _ptr(pos), _ptr(dzc), pos.shape[0], float(eps_scale), _ptr(cnt)))
        return dict(out=int(cnt[0]), inside=int(cnt[1]), unsure=int(cnt[2]), wrong=int(cnt[3]))

    def accumulate_audit(self, eps_scale: float = 1.0) -> dict:
        """fast vs exact path over every sub-particle of the grids on the device (nothing deposited)."""
        cnt = np.zeros(4, np.uint64)
        self._check(self.lib.gh_cuda_accumulate_audit(self._ctx, float(eps_scale), _ptr(cnt)))
        return dict(out=int(cnt[0]), inside=int(cnt[1]), unsure=int(cnt[2]), wrong=int(cnt[3]))

    # -- point sources ---------------------------------------------------------------
    def get_point_sources(self, ps) -> int:
        """src/grid_tools.c:24-101: Poisson-sample the sources of every cell from the Gaussian density (between
        create_d_and_vr_fields and get_HI).  `ps` is an abi.GhCudaPsourcesParams.  Returns the total over all ranks."""
        n = C.c_longlong()
        self._check(self.lib.gh_cuda_get_point_sources(self._ctx, C.byref(ps), C.byref(n)))
        return n.value

    def mk_psources_maps(self) -> np.ndarray:
        """src/pixelize.c:58-148 (after get_HI): this rank's shells of maps_PS, mK."""
        out = np.zeros((self.n_shells_here, self.npix), np.float32)
        self._check(self.lib.gh_cuda_mk_psources_maps(self._ctx, _ptr(out)))
        return out

    def download_point_sources(self):
        """(source counts, Poisson means) of this rank's slab, [nz_here][N][N]."""
        n = self.n_grid
        ns = np.zeros((self.nz_here, n, n), np.int32)
        lam = np.zeros((self.nz_here, n, n), np.float32)
        self._check(self.lib.gh_cuda_download_point_sources(self._ctx, _ptr(ns), _ptr(lam)))
        return ns, lam

    # -- JoinT ingestion -----------------------------------------------------------
    def jt_merge_maps(self, components, nside_out: int, scale=None) -> np.ndarray:
        """merge_maps (src/main_jt.c:98-211) for this rank's shells: `components` is the list of component stacks in
        the reference's order, each a float32 array [n_shells_here][npix] or None for the stack mk_T_maps / run just
        left on the device (the cosmological signal); the sum is degraded / upgraded to nside_out by he_udgrade
        (src/healpix_extra.c:318-385).  Returns [n_shells_here][12 nside_out^2]."""
        n = len(components)
        keep = [None if c is None else np.ascontiguousarray(c, dtype=np.float32) for c in components]
        for c in keep:
            if c is not None and c.shape != (self.n_shells_here, self.npix):
                raise ValueError(f"component stack must be [{self.n_shells_here}][{self.npix}], got {c.shape}")
        ptrs = (C.c_void_p * n)(*[None if c is None else c.ctypes.data for c in keep])
        sc = None if scale is None else (C.c_double * n)(*[float(v) for v in scale])
        out = np.empty((self.n_shells_here, 12 * nside_out * nside_out), np.float32)
        self._check(self.lib.gh_cuda_jt_merge_maps(self._ctx, n, ptrs, sc, int(nside_out), _ptr(out)))
        return out

    def udgrade(self, maps, nside_out: int, nest: bool = False) -> np.ndarray:
        """he_udgrade (src/healpix_extra.c:318-385) of one map or a stack of maps held in host memory."""
        m = np.ascontiguousarray(maps, dtype=np.float32)
        stack = m.reshape(1, -1) if m.ndim == 1 else m
        nside_in = int(round((stack.shape[1] / 12) ** 0.5))
        out = np.empty((stack.shape[0], 12 * nside_out * nside_out), np.float32)
        self._check(self.lib.gh_cuda_udgrade(self._ctx, _ptr(stack), nside_in, _ptr(out), int(nside_out), int(nest), stack.shape[0]))
        return out[0] if m.ndim == 1 else out

    def nest_ring(self, nside: int, pix, to_ring: bool) -> np.ndarray:
        """chealpix nest2ring (to_ring) / ring2nest through the device code of the two calls above."""
        a = np.ascontiguousarray(pix, dtype=np.int64)
        out = np.empty_like(a)
        self._check(self.lib.gh_cuda_nest_ring(self._ctx, int(nside), _ptr(a), _ptr(out), a.size, int(to_ring)))
        return out

    def stage_times(self) -> dict:
        ms = (C.c_double * len(STAGE_NAMES))()
        self._check(self.lib.gh_cuda_stage_times(self._ctx, ms))
        return dict(zip(STAGE_NAMES, list(ms)))

    def fft_pass_times(self) -> tuple[float, float]:
        """ms of the density / potential z pass incl. the fused transpose (-1: GH_TIME_FFT_PASSES was not set)."""
        ms = (C.c_double * 2)()
        self._check(self.lib.gh_cuda_fft_pass_times(self._ctx, ms))
        return float(ms[0]), float(ms[1])

    def kernel_launches(self) -> int:
        return int(self.lib.gh_cuda_kernel_launches(self._ctx))

    def stream_handle(self) -> int:
        return int(self.lib.gh_cuda_stream(self._ctx) or 0)

    # -- internals ---------------------------------------------------------------------------------
    def _host_maps(self, slot: int = 0) -> np.ndarray:
        n = max(self.n_shells_here, 1) * self.npix
        if self._pinned is None:
            self._pinned = {}
        if slot not in self._pinned:
            p = C.c_void_p()
            self._check(self.lib.gh_cuda_host_alloc(C.byref(p), n * 4))
            arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,))
            self._pinned[slot] = (p, arr)
        return self._pinned[slot][1][: self.n_shells_here * self.npix].reshape(self.n_shells_here, self.npix)

    def _check(self, rc: int):
        if rc != 0:
            raise GetHIError(self.lib.gh_cuda_last_error().decode(errors="replace"))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.end_fftw()

    def __del__(self):
        try:
            self.end_fftw()
        except Exception:
            pass
