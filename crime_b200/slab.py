"""Slab-ownership arithmetic shared by the host side and the tests (pure integers, no compute).

Mirrors what gh_cuda_create fixes per rank (crime_b200/csrc/gh_api.cu) and the reference's FFTW-MPI slab
decomposition (src/fourier.c:141-148): rank r of P owns z planes [r*N/P,(r+1)*N/P) in real space, the ky
rows of the same range in k-space (before the one transpose), and ceil(n_nu/P) consecutive shells of the
reduced map stack.
"""
from __future__ import annotations


def slab_bounds(n_grid: int, nranks: int, rank: int) -> tuple[int, int]:
    """(nz_here, iz0_here)."""
    if nranks < 1 or nranks & (nranks - 1) or n_grid % nranks or n_grid // nranks < 2:
        raise ValueError(f"n_grid={n_grid} cannot be split into {nranks} slabs")
    nz = n_grid // nranks
    return nz, rank * nz


def shell_bounds(n_nu: int, nranks: int, rank: int) -> tuple[int, int, int]:
    """(n_shells_here, shell0_here, n_nu_padded)."""
    per = -(-n_nu // nranks)
    s0 = rank * per
    return max(0, min(per, n_nu - s0)), s0, per * nranks


def transpose_chunk(n_grid: int, nranks: int) -> int:
    """complex elements each rank sends to each peer in the one all-to-all of a field."""
    nz = n_grid // nranks
    return nz * nz * (n_grid // 2 + 1)


def received_index(n_grid: int, nranks: int, z_local: int, ky: int, kx: int) -> int:
    """offset of mode (z_local, ky, kx) in the receive buffer [q][z_local][ky_local][kx]."""
    nyl = n_grid // nranks
    nh = n_grid // 2 + 1
    q, j = divmod(ky, nyl)
    return q * transpose_chunk(n_grid, nranks) + (z_local * nyl + j) * nh + kx


def map_plane_ranges(params, nranks: int) -> list[tuple[int, int]]:
    """[(first, last+1)] plane range each rank accumulates in mk_T_maps (gh_cuda_map_plane_bounds): equal modelled
    cost instead of equal plane counts.  Host-only; needs the built library but no GPU."""
    import ctypes as C
    from . import abi
    lib = abi.load_library()
    b = (C.c_int * (nranks + 1))()
    if lib.gh_cuda_map_plane_bounds(C.byref(params), nranks, b):
        raise RuntimeError(lib.gh_cuda_last_error().decode())
    return [(b[r], b[r + 1]) for r in range(nranks)]
