"""N>1 host logic on CPU: two gloo ranks run the slab-decomposed path with the oracle's kernels and must
reproduce the single-rank oracle (see tests/dist_cpu_worker.py)."""
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


import pytest


# obs_z: observer's z as a fraction of the box.  Off centre, the equal-cost plane ranges of the map accumulation
# differ from the slabs, so a rank also accumulates planes it pulled from the other.
@pytest.mark.parametrize("obs_z", [0.5, 0.2])
def test_two_rank_slab_decomposition_equals_single_rank(tmp_path, oracle, tables_nu64, obs_z):
    from crime_b200.gethi import params_from_tables
    n, n_side, world = 16, 8, 2
    procs = [subprocess.Popen([sys.executable, str(ROOT / "tests" / "dist_cpu_worker.py"), str(r), str(world), str(n), str(n_side),
                               str(tmp_path), str(obs_z)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    p = params_from_tables(tables_nu64, n_grid=n, n_side=n_side, seed=99)
    p.pos_obs[2] = obs_z * p.l_box
    ref = oracle.run(p)
    ranges = [tuple(np.load(tmp_path / f"rank{r}.npz")["plane_range"]) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n and ranges[0][1] == ranges[1][0]
    assert (ranges[0][1] != n // world) == (obs_z != 0.5)
    nz = n // world
    got = 0
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        sl = slice(int(d["iz0"]), int(d["iz0"]) + nz)
        assert np.array_equal(d["dens"][:, :, :n], ref["dens"][sl, :, :n])     # same arithmetic, just re-ordered ownership
        assert np.array_equal(d["rvel"][:, :, :n], ref["rvel"][sl, :, :n])
        assert np.array_equal(d["mass"][:, :, :n], ref["mass"][sl, :, :n])
        assert abs(float(d["sigma2"]) - ref["sigma2"]) <= 1e-12 * ref["sigma2"]
        s0 = int(d["s0"])
        m = d["maps"]
        rm = ref["maps"][s0:s0 + m.shape[0]]
        assert np.array_equal(m != 0, rm != 0)
        sel = rm != 0
        assert np.abs(m[sel] / rm[sel] - 1).max() < 2e-6   # float accumulation order across ranks
        got += m.shape[0]
    assert got == p.n_nu
