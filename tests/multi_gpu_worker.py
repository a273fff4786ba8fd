"""Worker for tests/test_gpu_multi.py: run under torchrun with one rank per GPU.

Every rank joins the slab-decomposed run (k-gen -> z FFT -> NCCL all-to-all -> y/x FFT -> halo exchange ->
velocity -> variance all-reduce -> get_HI -> maps -> reduce-scatter by shell); rank 0 also runs the same
problem alone on its GPU and checks that the decomposed result is the same."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from crime_b200 import GetHI, abi, params_from_tables  # noqa: E402
from crime_b200.abi import GRID_DENS, GRID_RVEL, GRID_VPOT  # noqa: E402


def main():
    n_grid, n_side = int(sys.argv[1]), int(sys.argv[2])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    raw = [None]
    if rank == 0:
        import ctypes as C
        buf = C.create_string_buffer(abi.GH_CUDA_UNIQUE_ID_BYTES)
        lib = abi.load_library()
        assert lib.gh_cuda_get_unique_id(buf) == 0, lib.gh_cuda_last_error()
        raw[0] = bytes(buf.raw)
    dist.broadcast_object_list(raw, 0)
    tables = dict(np.load(ROOT / "tests" / "golden" / "ref_tables_nu150.npz"))
    p = params_from_tables(tables, n_grid=n_grid, n_side=n_side, seed=31337)
    g = GetHI(p, rank=rank, nranks=world, unique_id=raw[0], device=local)
    assert g.nz_here == n_grid // world and g.iz0_here == rank * g.nz_here
    g.generate_k()
    dk, vk = g.download_delta_k()          # before the in-place FFTs overwrite it
    s2 = g.create_d_and_vr_fields()
    slabs = {k: g.download_grid(w) for k, w in (("dens", GRID_DENS), ("vpot", GRID_VPOT), ("rvel", GRID_RVEL))}
    g.get_HI()
    mass = g.download_grid(GRID_DENS)
    maps = g.mk_T_maps().copy()
    # the one-call path (gh_cuda_run: its own stage order, fused passes where enabled) must give the same maps
    maps_run = g.run().copy()
    run_ok = bool(np.array_equal(maps_run != 0, maps != 0))
    if run_ok and (maps != 0).any():
        run_ok = bool(np.abs(maps_run[maps != 0] / maps[maps != 0] - 1).max() < 1e-5)
    shells = (g.shell0_here, g.n_shells_here)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(dict(slabs=slabs, dk=dk, mass=mass, maps=maps, shells=shells, s2=s2, iz0=g.iz0_here, run_ok=run_ok),
                       gathered, 0)
    g.end_fftw()
    ok = True
    if rank == 0:
        with GetHI(p, device=local) as one:
            one.generate_k()
            dk_one, _ = one.download_delta_k()
            s2_one = one.create_d_and_vr_fields()
            ref = {k: one.download_grid(w) for k, w in (("dens", GRID_DENS), ("vpot", GRID_VPOT), ("rvel", GRID_RVEL))}
            one.get_HI()
            mass_one = one.download_grid(GRID_DENS)
            maps_one = one.mk_T_maps().copy()
        n = n_grid
        nz = n // world
        dk_all = np.zeros_like(dk_one)
        for r, part in enumerate(gathered):
            sl = slice(part["iz0"], part["iz0"] + nz)
            for k in ("dens", "vpot", "rvel"):
                if not np.array_equal(part["slabs"][k][:, :, :n], ref[k][sl, :, :n]):
                    err = np.abs(part["slabs"][k][:, :, :n] - ref[k][sl, :, :n]).max() / ref[k][:, :, :n].std()
                    print(f"rank {r} {k}: slab differs from the single-GPU field, max err/rms {err:.3e}")
                    ok = ok and err < 1e-6
            if not np.array_equal(part["mass"][:, :, :n], mass_one[sl, :, :n]):
                print(f"rank {r}: HI mass slab differs"); ok = False
            dk_all[:, r * nz:(r + 1) * nz] = part["dk"][:, r * nz:(r + 1) * nz]
            s0, ns = part["shells"]
            a, b = part["maps"], maps_one[s0:s0 + ns]
            if not np.array_equal(a != 0, b != 0):
                print(f"rank {r}: lit pixels differ"); ok = False
            nzm = b != 0
            if nzm.any() and np.abs(a[nzm] / b[nzm] - 1).max() > 1e-5:
                print(f"rank {r}: map values differ {np.abs(a[nzm] / b[nzm] - 1).max():.3e}"); ok = False
            if not part["run_ok"]:
                print(f"rank {r}: gh_cuda_run's maps differ from the staged calls'"); ok = False
            if abs(part["s2"] - s2_one) > 1e-12 * s2_one:
                print(f"rank {r}: sigma2 {part['s2']} vs {s2_one}"); ok = False
        # the k-space realisation does not depend on the number of slabs
        if not np.array_equal(dk_all, dk_one):
            print("k-space realisation depends on the decomposition"); ok = False
        assert sum(part["shells"][1] for part in gathered) == p.n_nu
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
