"""Grids whose size is not a power of two (the reference's FFTW takes any n_grid, /root/reference/src/fourier.c:85-99):
the general-length FFT passes of crime_b200/csrc/gh_fft.cu (fft_generic_*_kernel over gh_fft_generic.cuh) inside a whole
realisation, every intermediate field against the oracle -- the same assertions as test_whole_path_against_oracle, plus the
FFT alone against numpy and gh_cuda_run against the staged calls.  Both pass variants (one thread per butterfly, the
default; one thread per output element, GH_FFT_GENERIC_SLOW=1) were first run on a B200 at the very end of round 2:
profiles/r2/generic_grid_hw_*.log.  Each case runs in a process of its own under a timeout and the file sorts last, so that
this newest code path cannot take the rest of the suite with it."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


# 48 = 4^2 3, 80 = 4^2 5, 96 = 4^2 2 3: whole 16-cell bricks of the map kernel; 40 = 4 2 5 and 56 = 4 2 7 end in a partial brick;
# 44 = 4 11: a radix without a butterfly of its own (one thread per output element); 384 = 4^3 2 3: a production-sized grid
@pytest.mark.parametrize("n_grid,n_side,variant", [(48, 16, "butterfly"), (80, 32, "butterfly"), (96, 32, "butterfly"), (40, 16, "butterfly"),
                                                   (56, 16, "butterfly"), (44, 16, "butterfly"), (48, 16, "per_output"), (40, 16, "per_output"),
                                                   (384, 128, "butterfly")])
def test_non_power_of_two_grid_against_oracle(n_grid, n_side, variant):
    cmd = [sys.executable, str(ROOT / "tests" / "generic_grid_worker.py"), str(n_grid), str(n_side)]
    env = dict(os.environ)
    env.pop("GH_FFT_GENERIC_SLOW", None)
    if variant == "per_output":
        env["GH_FFT_GENERIC_SLOW"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    out = ROOT / "gpurun_out" / "generic_grid"
    out.mkdir(parents=True, exist_ok=True)
    (out / f"n{n_grid}_{variant}.log").write_text(r.stdout[-4000:] + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0 and "GENERIC_GRID_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
