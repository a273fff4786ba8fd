"""Grids whose size is not a power of two (the reference's FFTW takes any n_grid, /root/reference/src/fourier.c:85-99):
the general-length FFT passes of crime_b200/csrc/gh_fft.cu (fft_generic_*_kernel over gh_fft_generic.cuh) inside a whole
realisation, every intermediate field against the oracle -- the same assertions as test_whole_path_against_oracle.

These kernels were written after this round's GPU budget was spent: their arithmetic, thread mapping, shared-memory
addressing and launch geometry are verified on the CPU by executing the very header the kernels include
(tests/test_fft_generic_cpu.py), but the first hardware run is whoever runs this file.  Hence: each case runs in a process
of its own under a timeout (a faulting or hanging kernel cannot take the suite with it), the file sorts last, and the cases
are xfail(strict=False) -- XPASS means the path works on the device, xfail means it does not yet; neither hides or breaks
the validated tests before it."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


# 48 = 4^2 3, 80 = 4^2 5, 96 = 4^2 2 3: whole 16-cell bricks of the map kernel; 40 = 4 2 5 and 56 = 4 2 7 end in a partial brick;
# 44 = 4 11: a radix the plan has no special case for
@pytest.mark.xfail(strict=False, reason="general-length FFT kernels: verified on the CPU only so far (no GPU budget left when written)")
@pytest.mark.parametrize("n_grid,n_side", [(48, 16), (80, 32), (96, 32), (40, 16), (56, 16), (44, 16)])
def test_non_power_of_two_grid_against_oracle(n_grid, n_side):
    cmd = [sys.executable, str(ROOT / "tests" / "generic_grid_worker.py"), str(n_grid), str(n_side)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    out = ROOT / "gpurun_out" / "generic_grid"
    out.mkdir(parents=True, exist_ok=True)
    (out / f"n{n_grid}.log").write_text(r.stdout[-4000:] + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0 and "GENERIC_GRID_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
