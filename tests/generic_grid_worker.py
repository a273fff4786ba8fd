"""Worker of tests/test_zz_gpu_generic_grid.py: one GetHI realisation on a grid whose size is not a power of two
(the general-length FFT passes of gh_fft.cu), every intermediate field against the oracle, in a process of its own so
that a faulting kernel cannot take the rest of the GPU suite with it.  Prints one GENERIC_GRID_OK line with the numbers."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

TOL = 1e-5


def main(n, n_side):
    from conftest import field_err
    from crime_b200 import GetHI, params_from_tables
    from crime_b200.abi import GRID_DENS, GRID_RVEL, GRID_VPOT
    from oracle.binding import Oracle
    orc = Oracle()
    tabs = dict(np.load(ROOT / "tests" / "golden" / "ref_tables_nu64.npz"))
    p = params_from_tables(tabs, n_grid=n, n_side=n_side, seed=4242)
    dk_o, vk_o = orc.kgen_philox(p)
    res = {"n_grid": n, "n_side": n_side}
    with GetHI(p) as g:
        g.generate_k()
        dk, vk = g.download_delta_k()
        m = np.abs(dk_o) > 0
        res["kgen"] = float((np.abs(dk - dk_o)[m] / np.abs(dk_o)[m]).max())
        assert res["kgen"] < TOL, res
        # the FFT alone against numpy's c2r on the device's own k-space
        g.set_delta_k(dk_o, vk_o)
        s2 = g.create_d_and_vr_fields()
        dens = g.download_grid(GRID_DENS)
        vpot = g.download_grid(GRID_VPOT)
        rvel = g.download_grid(GRID_RVEL)
        want = np.fft.irfftn(dk_o.astype(np.complex128), s=(n, n, n), axes=(0, 1, 2)) * float(n) ** 3 * (np.sqrt(2 * np.pi) / p.l_box) ** 3
        res["fft_vs_numpy"] = field_err(dens[:, :, :n], want)
        assert res["fft_vs_numpy"] < TOL, res
        o = orc.run(p, dk_o, vk_o)
        res["dens"] = field_err(dens[:, :, :n], o["dens"][:, :, :n])
        res["vpot"] = field_err(vpot[:, :, :n], o["vpot"][:, :, :n])
        res["rvel"] = field_err(rvel[:, :, :n], o["rvel"][:, :, :n])
        res["sigma2"] = abs(s2 - o["sigma2"]) / s2
        assert res["dens"] < TOL and res["vpot"] < TOL and res["rvel"] < 20 * TOL and res["sigma2"] < TOL, res
        g.upload_grid(GRID_DENS, o["dens"])
        g.upload_grid(GRID_RVEL, o["rvel"])
        g.set_sigma2_gauss(o["sigma2"])
        g.get_HI()
        mass = g.download_grid(GRID_DENS)
        dz = g.download_grid(GRID_RVEL)
        res["mass"] = float(np.abs(mass[:, :, :n] / o["mass"][:, :, :n] - 1).max())
        res["dz"] = field_err(dz[:, :, :n], o["dz"][:, :, :n])
        assert res["mass"] < TOL and res["dz"] < TOL, res
        g.upload_grid(GRID_DENS, o["mass"])
        g.upload_grid(GRID_RVEL, o["dz"])
        maps = g.mk_T_maps().copy()
        ref = o["maps"]
        res["lit_pixels_equal"] = bool(np.array_equal(maps != 0, ref != 0))
        nz = ref != 0
        res["lit_pixels"] = int(nz.sum())
        res["maps"] = float(np.abs(maps[nz] / ref[nz] - 1).max()) if res["lit_pixels_equal"] else None
        assert res["lit_pixels_equal"] and res["maps"] < TOL, res
        # the whole run in one call (fused velocity + get_HI) equals the staged calls on the device's own realisation
        g.clear_delta_k()
        maps_run = g.run().copy()
        g.create_d_and_vr_fields()
        g.get_HI()
        maps_staged = g.mk_T_maps().copy()
        res["run_equals_staged_lit"] = bool(np.array_equal(maps_run != 0, maps_staged != 0))
        nz = maps_staged != 0
        res["run_vs_staged"] = float(np.abs(maps_run[nz] / maps_staged[nz] - 1).max())
        assert res["run_equals_staged_lit"] and res["run_vs_staged"] < TOL, res
        res["kernel_launches"] = int(g.kernel_launches())
    print("GENERIC_GRID_OK " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if ":" in sys.argv[1]:  # several cases in one process: n_grid:n_side ...
        bad = 0
        for spec in sys.argv[1:]:
            n, ns = (int(v) for v in spec.split(":"))
            try:
                main(n, ns)
            except Exception as exc:  # noqa: BLE001 -- report and go on to the next size
                bad += 1
                print(f"GENERIC_GRID_FAILED n_grid={n}: {type(exc).__name__}: {str(exc)[:600]}", flush=True)
        sys.exit(1 if bad else 0)
    main(int(sys.argv[1]), int(sys.argv[2]))
