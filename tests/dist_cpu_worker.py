"""World-size-2 (gloo, CPU) restatement of the slab-decomposed GetHI path with the oracle's kernels.

Each rank follows exactly the data movement libgh_cuda.so performs on GPUs (crime_b200/csrc/gh_fft.cu,
gh_fields.cu, gh_api.cu) -- ky-distributed k-space, local z transform, ONE all-to-all per field, local y/x
transform, ring halo exchange of the potential, all-reduce of two doubles, reduce-scatter of the map stack
by padded shells -- with crime_b200.slab supplying the ownership arithmetic.  The result must equal the
single-rank oracle.  Launched by tests/test_dist_cpu.py."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from crime_b200 import slab  # noqa: E402
from crime_b200.gethi import params_from_tables  # noqa: E402
from oracle.binding import Oracle, Slab, _ptr  # noqa: E402
import ctypes as C  # noqa: E402


def run(rank: int, world: int, n: int, n_side: int, out_dir: str, obs_z: float = 0.5):
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    tables = dict(np.load(ROOT / "tests" / "golden" / "ref_tables_nu64.npz"))
    p = params_from_tables(tables, n_grid=n, n_side=n_side, seed=99)
    p.pos_obs[2] = obs_z * p.l_box   # an off-centre observer makes the equal-cost plane ranges differ from the slabs
    nh = n // 2 + 1
    nz, iz0 = slab.slab_bounds(n, world, rank)
    # (1) k-space for this rank's ky rows, layout [kz][ky_local][kx]
    dk, vk = orc.kgen_philox(p, iz0, nz)
    fields = []
    for k in (dk, vk):
        k = np.ascontiguousarray(k)
        orc.lib.oracle_fft_axis0(n, nz, nh, _ptr(k))  # (2) z transform, local
        # (3) the one transpose: block q = kz in rank q's slab
        send = torch.from_numpy(k.view(np.float32).reshape(world, -1).copy())
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        chunk = slab.transpose_chunk(n, world)
        rbuf = recv.numpy().view(np.complex64).reshape(-1)
        assert rbuf.size == world * chunk
        plane = np.empty((nz, n, nh), np.complex64)
        for z in range(nz):  # received layout [q][z_local][ky_local][kx] -> [z_local][ky][kx]
            for ky in range(n):
                o = slab.received_index(n, world, z, ky, 0)
                plane[z, ky] = rbuf[o:o + nh]
        orc.lib.oracle_fft_axis1_c2r_axis2(n, nz, _ptr(plane))  # (4) y transform + x c2r
        fields.append(plane.view(np.float32).reshape(nz, n, 2 * nh))
    dens, vpot = fields
    orc.lib.oracle_normalize(C.byref(p), Slab(nz, iz0), _ptr(dens), _ptr(vpot))
    # (5) halo exchange on the ring (src/fourier.c:415-424)
    right, left = (rank + 1) % world, (rank - 1) % world
    lo, hi = torch.empty(n, 2 * nh), torch.empty(n, 2 * nh)
    reqs = [dist.isend(torch.from_numpy(vpot[-1].copy()), right), dist.irecv(lo, left),
            dist.isend(torch.from_numpy(vpot[0].copy()), left), dist.irecv(hi, right)]
    for r in reqs:
        r.wait()
    rvel = orc.radial_velocity(p, vpot, iz0, lo.numpy(), hi.numpy())
    # (6) variance: all-reduce of two doubles
    m, s2 = orc.sigma_partial(p, dens, iz0)
    t = torch.tensor([m, s2], dtype=torch.float64)
    dist.all_reduce(t)
    sigma2 = float(t[1] - t[0] * t[0])
    mass, dz = orc.get_HI(p, sigma2, dens, rvel, iz0)
    # (7) maps: this rank accumulates the plane range the library's cost model gives it
    # (gh_cuda_map_plane_bounds); planes outside its slab are pulled from their owner (on GPUs: peer copies over
    # NVLink, here: an all-gather of the slabs); full per-rank stack, reduce-scatter over padded shells
    nsh, s0, npad = slab.shell_bounds(p.n_nu, world, rank)
    npix = 12 * n_side * n_side
    stack = np.zeros((npad, npix), np.float32)
    lo_p, hi_p = slab.map_plane_ranges(p, world)[rank]
    gm = [torch.empty(nz, n, 2 * nh) for _ in range(world)]
    gz = [torch.empty(nz, n, 2 * nh) for _ in range(world)]
    dist.all_gather(gm, torch.from_numpy(np.ascontiguousarray(mass)))
    dist.all_gather(gz, torch.from_numpy(np.ascontiguousarray(dz)))
    all_m, all_z = torch.cat(gm).numpy(), torch.cat(gz).numpy()
    if hi_p > lo_p:
        orc.accumulate_maps(p, np.ascontiguousarray(all_m[lo_p:hi_p]), np.ascontiguousarray(all_z[lo_p:hi_p]), lo_p, stack[:p.n_nu])
    mine = torch.empty(npad // world, npix)
    dist.reduce_scatter_tensor(mine, torch.from_numpy(stack))
    mine = mine.numpy()[:nsh]
    pref = orc.shell_prefactors(p)[s0:s0 + nsh]
    mine = (mine.astype(np.float64) * pref[:, None]).astype(np.float32)
    np.savez(f"{out_dir}/rank{rank}.npz", dens=dens, rvel=rvel, mass=mass, maps=mine, sigma2=sigma2, iz0=iz0, s0=s0,
             plane_range=np.array([lo_p, hi_p]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5],
        float(sys.argv[6]) if len(sys.argv) > 6 else 0.5)
