"""One or a few realisations through the C-ABI, for profiler captures:  python tools/run_once.py [n_grid n_side n_nu [steps]]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import load_tables  # noqa: E402
from crime_b200.gethi import GetHI, params_from_tables  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
n, ns, nu, steps = (a + [512, 256, 64, 2][len(a):])[:4]
p = params_from_tables(load_tables(nu), n_grid=n, n_side=ns, seed=1001)
with GetHI(p) as g:
    for _ in range(steps):
        g.run(to_host=False)
    g.synchronize()
    print({k: round(v, 4) for k, v in g.stage_times().items()})
