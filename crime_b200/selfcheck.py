"""Slab-decomposed run == single-GPU run, as a callable.

Used by tests/multi_gpu_worker.py (the multi-GPU parity tests) and by bench.py --gpus N, which runs it on a small
grid before timing and prints the outcome as the `parity` object of its JSON line, so that a scaling record carries
its own evidence that the path it timed gives the single-GPU (and hence, through the single-GPU parity tests, the
oracle's) results.

Two modes:
  full : every rank's slabs and shells are gathered on rank 0 and compared element by element with a run of the same
         parameters on rank 0's GPU alone (fields must be bit-identical, maps equal to float-atomic order);
  hash : for grids too large to move (2048^3): per-slab position-weighted checksums of the five grids
         (gh_cuda_grid_checksum) against the single-GPU run's checksums of the same plane ranges; maps are still
         compared element-wise, shell by shell.

Everything here goes through the C-ABI (crime_b200.GetHI); nothing touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import zlib

import numpy as np

from . import abi
from .abi import GRID_DENS, GRID_RVEL, GRID_VPOT
from .gethi import GetHI

FIELDS = (("dens", GRID_DENS), ("vpot", GRID_VPOT), ("rvel", GRID_RVEL))


def make_unique_id(dist, rank: int) -> bytes:
    """Rank 0 asks the library for an NCCL id; everybody receives it through the launcher's process group."""
    raw = [None]
    if rank == 0:
        buf = C.create_string_buffer(abi.GH_CUDA_UNIQUE_ID_BYTES)
        lib = abi.load_library()
        if lib.gh_cuda_get_unique_id(buf):
            raise RuntimeError(lib.gh_cuda_last_error().decode())
        raw[0] = bytes(buf.raw)
    dist.broadcast_object_list(raw, 0)
    return raw[0]


def _maps_close(a: np.ndarray, b: np.ndarray):
    """(lit pixel sets equal, max relative difference on the lit pixels)."""
    same = bool(np.array_equal(a != 0, b != 0))
    nz = b != 0
    rel = float(np.abs(a[nz] / b[nz] - 1).max()) if same and nz.any() else (0.0 if same else float("inf"))
    return same, rel


def decomposed_vs_single(dist, params, rank: int, world: int, device: int, mode: str = "full", verbose: bool = False) -> dict | None:
    """Run `params` slab-decomposed over the `world` ranks of the initialised process group `dist` and alone on
    rank 0's GPU; returns the comparison on rank 0 (None elsewhere).  Collective: every rank must call it."""
    n = params.n_grid
    uid = make_unique_id(dist, rank)
    g = GetHI(params, rank=rank, nranks=world, unique_id=uid, device=device)
    nz = g.nz_here
    assert nz == n // world and g.iz0_here == rank * nz
    g.generate_k()
    part = {"iz0": g.iz0_here}
    if mode == "full":
        part["dk"] = g.download_delta_k()[0][:, rank * nz:(rank + 1) * nz].copy()   # before the in-place FFTs overwrite it
    part["s2"] = g.create_d_and_vr_fields()
    if mode == "full":
        part["slabs"] = {k: g.download_grid(w)[:, :, :n].copy() for k, w in FIELDS}
    else:
        part["sums"] = {k: g.grid_checksum(w) for k, w in FIELDS}
    g.get_HI()
    if mode == "full":
        part["mass"] = g.download_grid(GRID_DENS)[:, :, :n].copy()
        part["dz"] = g.download_grid(GRID_RVEL)[:, :, :n].copy()
    else:
        part["sums"]["mass"] = g.grid_checksum(GRID_DENS)
        part["sums"]["dz"] = g.grid_checksum(GRID_RVEL)
    maps = g.mk_T_maps().copy()
    # the one-call path (gh_cuda_run: its own stage order, fused passes where enabled) must give the same maps
    run_same, run_rel = _maps_close(g.run().copy(), maps)
    part.update(shells=(g.shell0_here, g.n_shells_here), run_ok=bool(run_same and run_rel < 1e-5))
    if mode == "full":
        part["maps"] = maps
    g.end_fftw()
    gathered = [None] * world
    dist.all_gather_object(gathered, part)
    shm = f"/dev/shm/gh_selfcheck_{zlib.crc32(uid):08x}.npy"
    if rank != 0:
        if mode == "hash":  # the single-GPU maps arrive through /dev/shm: compare this rank's shells in place
            dist.barrier()
            ref = np.load(shm, mmap_mode="r")
            s0, ns = part["shells"]
            mine = _maps_close(maps, ref[s0:s0 + ns])
            del ref
            dist.all_gather_object([None] * world, mine)
        return None
    res = {"n_grid": n, "n_side": int(params.n_side), "n_nu": int(params.n_nu), "ranks": world, "mode": mode}
    with GetHI(params, device=device) as one:
        one.generate_k()
        dk_one = one.download_delta_k()[0] if mode == "full" else None
        s2_one = one.create_d_and_vr_fields()
        if mode == "full":
            ref = {k: one.download_grid(w)[:, :, :n] for k, w in FIELDS}
        else:
            ref_sums = [{k: one.grid_checksum(w, r * nz, nz) for k, w in FIELDS} for r in range(world)]
        one.get_HI()
        if mode == "full":
            ref["mass"] = one.download_grid(GRID_DENS)[:, :, :n]
            ref["dz"] = one.download_grid(GRID_RVEL)[:, :, :n]
        else:
            for r in range(world):
                ref_sums[r]["mass"] = one.grid_checksum(GRID_DENS, r * nz, nz)
                ref_sums[r]["dz"] = one.grid_checksum(GRID_RVEL, r * nz, nz)
        maps_one = one.mk_T_maps().copy()
    map_cmp = None
    if mode == "hash":
        import os
        np.save(shm, maps_one)
        dist.barrier()
        s0, ns = part["shells"]
        map_cmp = [None] * world
        dist.all_gather_object(map_cmp, _maps_close(maps, maps_one[s0:s0 + ns]))
        os.unlink(shm)
    fields_ok, lit_ok, max_rel, k_ok, s2_ok, run_ok, worst = True, True, 0.0, True, True, True, 0.0
    for r, p in enumerate(gathered):
        sl = slice(p["iz0"], p["iz0"] + nz)
        if mode == "full":
            for k in ("dens", "vpot", "rvel", "mass", "dz"):
                a = p["slabs"][k] if k in p.get("slabs", {}) else p[k]
                if not np.array_equal(a, ref[k][sl]):
                    fields_ok = False
                    err = float(np.abs(a - ref[k][sl]).max() / ref[k].std())
                    worst = max(worst, err)
                    if verbose:
                        print(f"rank {r} {k}: slab differs from the single-GPU field, max err/rms {err:.3e}")
            if not np.array_equal(p["dk"], dk_one[:, r * nz:(r + 1) * nz]):
                k_ok = False
        else:
            for k, v in p["sums"].items():
                if v != ref_sums[r][k]:
                    fields_ok = False
                    if verbose:
                        print(f"rank {r} {k}: checksum {v:#x} vs single-GPU {ref_sums[r][k]:#x}")
        s0, ns = p["shells"]
        same, rel = _maps_close(p["maps"], maps_one[s0:s0 + ns]) if mode == "full" else map_cmp[r]
        lit_ok = lit_ok and same
        max_rel = max(max_rel, rel)
        run_ok = run_ok and p["run_ok"]
        s2_ok = s2_ok and abs(p["s2"] - s2_one) <= 1e-12 * s2_one
        if verbose and not (same and rel < 1e-5 and p["run_ok"]):
            print(f"rank {r}: lit pixels equal {same}, max rel {rel:.3e}, run()==staged {p['run_ok']}")
    shells_ok = sum(p["shells"][1] for p in gathered) == params.n_nu
    res.update(fields_bit_identical=bool(fields_ok), lit_pixels_equal=bool(lit_ok), max_rel=float(max_rel),
               kspace_independent_of_slabs=bool(k_ok) if mode == "full" else None, sigma2_equal=bool(s2_ok),
               run_equals_staged_calls=bool(run_ok), shells_partitioned=bool(shells_ok),
               worst_field_err_over_rms=float(worst))
    res["ok"] = bool(fields_ok and lit_ok and max_rel < 1e-5 and k_ok and s2_ok and run_ok and shells_ok)
    return res
