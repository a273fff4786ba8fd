"""CPU check of the general-length FFT (crime_b200/csrc/gh_fft_generic.cuh): the reference's FFTW takes any n_grid
(/root/reference/src/fourier.c:85-99), the tuned sm_100a kernels exist for powers of two, and every other even n_grid
runs through the CTA phase functions of this header.  The very header the kernels include is compiled for the host
(tests/native/fft_generic_host.cpp) and executed block by block, phase by phase, thread by thread with the launcher's own
geometry -- whole cubes against numpy's c2r for small grids, single passes for the lengths a production grid would have
(384 ... 4094).  Guard zones around the emulated shared memory and the field catch out-of-range accesses, and running the
threads of a phase backwards must not change a bit (the phases are separated by __syncthreads() only).
The kernels themselves: tests/test_zz_gpu_generic_grid.py."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent / "native"


@pytest.fixture(scope="module")
def lib():
    so = HERE / "libfft_generic_host.so"
    src = HERE / "fft_generic_host.cpp"
    hdr = HERE.parents[1] / "crime_b200" / "csrc" / "gh_fft_generic.cuh"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-x", "c++", "-o", str(so), str(src)], check=True)
    return ctypes.CDLL(str(so))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(params=[0, 1], ids=["per_output", "per_butterfly"])
def fast(lib, request):
    """0: one thread per output element in every pass (gfft_pass); 1: one thread per butterfly for the radices 2, 3, 4, 5, 7
    (gfft_pass_small), the rest as before."""
    lib.gfft_host_set_fast(request.param)
    yield request.param
    lib.gfft_host_set_fast(0)


def plan_of(lib, n):
    f = np.zeros(16, np.int32)
    k = lib.gfft_host_plan(n, _ptr(f))
    return list(f[:k])


@pytest.mark.parametrize("n", [1, 2, 3, 4, 6, 12, 25, 48, 96, 97, 100, 384, 768, 1000, 1536, 2047, 2304, 3000, 4094, 4096])
def test_radix_plan(lib, n):
    f = plan_of(lib, n)
    assert int(np.prod(f, dtype=np.int64)) == n
    assert all(r == 4 or all(r % q for q in range(2, int(r ** 0.5) + 1)) for r in f), f  # 4 or prime
    assert f.count(2) <= 1


@pytest.mark.parametrize("n", [8, 10, 12, 20, 24, 36, 42, 48, 50, 66, 96, 100, 768, 1536, 2304, 3072, 4094])
def test_launch_geometry_fits_the_sm(lib, n):
    out = np.zeros(8, np.int64)
    assert lib.gfft_host_launch(n, n, _ptr(out)) == 0
    W, WR, pitch, smem_s, smem_r, bz, by, bx = (int(v) for v in out)
    assert 1 <= W <= 16 and 1 <= WR <= 16 and pitch % 2 == 1 and WR <= pitch <= WR + 1
    assert smem_s == 16 * n * W <= 200 * 1024 and smem_r == 16 * (n // 2) * pitch <= 200 * 1024
    nh = n // 2 + 1
    assert bz == -(-n * nh // W) and by == n * -(-nh // W) and bx == -(-n * n // WR)
    assert max(bz, by, bx) < 2 ** 31


def test_odd_or_oversized_lengths_are_refused(lib):
    out = np.zeros(8, np.int64)
    assert lib.gfft_host_launch(47, 47, _ptr(out)) != 0
    assert lib.gfft_host_launch(16384, 2, _ptr(out)) != 0


def _spectrum(rng, n, hermitian_planes):
    nh = n // 2 + 1
    x = (rng.standard_normal((n, n, nh)) + 1j * rng.standard_normal((n, n, nh))).astype(np.complex64)
    if hermitian_planes:
        # a genuine half-complex spectrum of a real field
        x = np.fft.rfftn(rng.standard_normal((n, n, n))).astype(np.complex64)
    return x


@pytest.mark.parametrize("n", [8, 10, 12, 20, 24, 36, 42, 48, 50, 66, 96, 100])
@pytest.mark.parametrize("hermitian", [True, False])
def test_whole_cube_against_numpy(lib, fast, n, hermitian):
    """The three launches of fft_field_generic on an n^3 grid = numpy's c2r (which, like FFTW's and like the tuned
    kernels, never reads Im of the kx = 0 and kx = n/2 planes' self-conjugate partners: non-Hermitian input is
    projected the same way), normalisation included."""
    rng = np.random.default_rng(n)
    x = _spectrum(rng, n, hermitian)
    norm = 0.37
    want = np.fft.irfftn(x.astype(np.complex128), s=(n, n, n), axes=(0, 1, 2)) * float(n) ** 3 * norm
    got = {}
    for reverse in (0, 1):
        buf = np.ascontiguousarray(x).view(np.float32).copy()
        assert lib.gfft_host_field(_ptr(buf), n, ctypes.c_double(norm), 256, reverse) == 0
        got[reverse] = buf.reshape(n, n, 2 * (n // 2 + 1))
    assert np.array_equal(got[0], got[1]), "a phase depends on the order of its threads"
    real = got[0][:, :, :n]
    err = np.abs(real - want).max() / want.std()
    assert err < 2e-6, err


@pytest.mark.parametrize("nthreads", [32, 96, 256, 1024])
def test_any_block_size(lib, fast, nthreads):
    n = 24
    rng = np.random.default_rng(5)
    x = _spectrum(rng, n, False)
    want = np.fft.irfftn(x.astype(np.complex128), s=(n, n, n), axes=(0, 1, 2)) * float(n) ** 3
    buf = np.ascontiguousarray(x).view(np.float32).copy()
    assert lib.gfft_host_field(_ptr(buf), n, ctypes.c_double(1.0), nthreads, 0) == 0
    real = buf.reshape(n, n, 2 * (n // 2 + 1))[:, :, :n]
    assert np.abs(real - want).max() / want.std() < 2e-6


@pytest.mark.parametrize("n", [384, 768, 1000, 1536, 2304, 3000, 3072, 4094])
def test_production_lengths_pass_by_pass(lib, fast, n):
    """Lengths a production grid would use (incl. 4094 = 2 * 23 * 89: large prime radices): one strided pass over a few
    tiles' worth of lines (the last tile ragged) and one x pass over a few tiles' worth of rows, with the tile widths
    the launcher picks for this length."""
    rng = np.random.default_rng(n)
    out = np.zeros(8, np.int64)
    assert lib.gfft_host_launch(n, n, _ptr(out)) == 0
    W, WR = int(out[0]), int(out[1])
    lines = 2 * W + max(1, W // 2)
    x = (rng.standard_normal((n, lines)) + 1j * rng.standard_normal((n, lines))).astype(np.complex64)
    want = np.fft.ifft(x.astype(np.complex128), axis=0) * n
    res = {}
    for reverse in (0, 1):
        buf = np.ascontiguousarray(x).view(np.float32).copy()
        assert lib.gfft_host_strided_lines(_ptr(buf), n, lines, 256, reverse) == 0
        res[reverse] = buf
    assert np.array_equal(res[0], res[1])
    got = res[0].view(np.complex64).reshape(n, lines)
    assert np.abs(got - want).max() / np.abs(want).std() < 3e-6

    nrows, nh = 2 * WR + 1, n // 2 + 1
    x = (rng.standard_normal((nrows, nh)) + 1j * rng.standard_normal((nrows, nh))).astype(np.complex64)
    want = np.fft.irfft(x.astype(np.complex128), n=n, axis=1) * n * 2.5
    buf = np.ascontiguousarray(x).view(np.float32).copy()
    assert lib.gfft_host_rows(_ptr(buf), n, ctypes.c_longlong(nrows), ctypes.c_double(2.5), 256, 0) == 0
    got = buf.reshape(nrows, 2 * nh)[:, :n]
    assert np.abs(got - want).max() / want.std() < 3e-6


@pytest.mark.parametrize("n", [18, 30, 42, 50, 70, 90, 98])
def test_odd_radix_butterflies_equal_the_per_output_passes_bit_for_bit(lib, n):
    """Radices 3, 5 and 7 do the same products in the same order either way (only 2 and 4 trade table look-ups of -1 and
    +-i, whose float sines are 1e-16 rather than 0, for exact butterflies).  The x pass of these grids transforms an odd
    half-length (9 ... 49): any difference between the two variants would be an indexing slip, not rounding."""
    assert all(r % 2 for r in plan_of(lib, n // 2))
    rng = np.random.default_rng(n)
    nrows, nh = 7, n // 2 + 1
    x = (rng.standard_normal((nrows, nh)) + 1j * rng.standard_normal((nrows, nh))).astype(np.complex64)
    out = {}
    for fast in (0, 1):
        lib.gfft_host_set_fast(fast)
        buf = np.ascontiguousarray(x).view(np.float32).copy()
        assert lib.gfft_host_rows(_ptr(buf), n, ctypes.c_longlong(nrows), ctypes.c_double(1.0), 128, 0) == 0
        out[fast] = buf
    lib.gfft_host_set_fast(0)
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32))
