"""Test infrastructure (it uses oracle/_ref as the checker, so it lives under tests/).  A few seconds on a GPU box, no torch /
pytest: (a) the FITS maps of a real ./GetHI run read back through the compiled
reference's he_read_healpix_map, (b) the general-length FFT passes against numpy at 192^3 and their time at 384^3 / 768^3
next to the tuned 512^3 kernels.  Every result is printed (flushed) as soon as it exists."""
import ctypes as C
import json
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from crime_b200 import GetHI, host, params_from_tables  # noqa: E402
from crime_b200.abi import GRID_DENS  # noqa: E402
from oracle.binding import Reference, write_nutable, write_param_file  # noqa: E402


def say(tag, **kw):
    print(tag + " " + json.dumps(kw), flush=True)


def fits_through_the_reference_reader():
    tmp = Path(tempfile.mkdtemp())
    write_nutable(tmp / "nu.txt", 16)
    write_param_file(tmp / "p.ini", n_grid=64, n_side=32, nutable=tmp / "nu.txt", pk_file=ROOT / "data" / "Pk_synth.dat",
                     prefix=tmp / "run", seed=2024)
    r = subprocess.run([str(host.HOST_EXE), str(tmp / "p.ini")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    rd = Reference().lib.he_read_healpix_map
    rd.argtypes = [C.c_char_p, C.POINTER(C.c_long), C.c_int]
    rd.restype = C.POINTER(C.c_float)
    same, lit = 0, 0
    for s in range(16):
        f = tmp / f"run_{s + 1:03d}.fits"
        m, hdr = host.read_healpix_map(f)
        ns = C.c_long(-1)
        ptr = rd(str(f).encode(), C.byref(ns), 0)
        ok = ns.value == 32 and np.array_equal(np.ctypeslib.as_array(ptr, shape=(12 * 32 * 32,)), m)
        same += int(ok)
        lit += int((m != 0).sum())
    say("FITS_REFERENCE_READER", files=16, identical=same, lit_pixels=lit)


def fft_case(tabs, n, check):
    p = params_from_tables(tabs, n_grid=n, n_side=16, seed=7)
    with GetHI(p) as g:
        g.generate_k()
        err = None
        if check:
            dk, _ = g.download_delta_k()
        g.fft_fields()
        g.synchronize()
        if check:
            dens = g.download_grid(GRID_DENS)[:, :, :n]
            want = np.fft.irfftn(dk.astype(np.complex128), s=(n, n, n), axes=(0, 1, 2)) * float(n) ** 3 * (np.sqrt(2 * np.pi) / p.l_box) ** 3
            err = float(np.abs(dens - want).max() / want.std())
        ms = []
        for _ in range(3):
            g.generate_k()
            g.fft_fields()
            g.synchronize()
            ms.append(g.stage_times()["fft"])
        s2, _ = g.sigma_dens()
    say("FFT", n_grid=n, tuned=(n & (n - 1)) == 0, err_vs_numpy=err, fft_ms_both_fields=round(min(ms), 4),
        gcells_per_s=round(n ** 3 / (min(ms) * 1e-3) / 1e9, 2), sigma2=s2)


if __name__ == "__main__":
    t0 = time.time()
    tabs = dict(np.load(ROOT / "tests" / "golden" / "ref_tables_nu64.npz"))
    steps = [fits_through_the_reference_reader, lambda: fft_case(tabs, 192, True), lambda: fft_case(tabs, 384, False),
             lambda: fft_case(tabs, 512, False), lambda: fft_case(tabs, 768, False), lambda: fft_case(tabs, 640, False)]
    if "--fft-only" in sys.argv:  # the general-length passes alone
        steps = [lambda: fft_case(tabs, 192, True), lambda: fft_case(tabs, 384, False), lambda: fft_case(tabs, 768, False),
                 lambda: fft_case(tabs, 640, False), lambda: fft_case(tabs, 1536, False)]
    for step in steps:
        try:
            step()
        except Exception as exc:  # noqa: BLE001
            say("FAILED", error=f"{type(exc).__name__}: {str(exc)[:500]}")
    say("DONE", seconds=round(time.time() - t0, 1))
