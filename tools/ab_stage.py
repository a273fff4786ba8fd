"""A/B harness for opt-in kernel variants (run on a GPU box): for each environment setting, the stage times of a
whole realisation, the map stage alone, the on-device audit of the fast path against the exact path, and the maps
compared with the default build's.

    python tools/ab_stage.py                      # the default build alone at 512^3
    python tools/ab_stage.py 1024 512 150 GH_NO_FUSE_VEL=1 GH_CUDA_LIB=crime_b200/csrc/variants/libgh_cuda_x.so
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def child(n, ns, nu):
    import numpy as np
    from bench import load_tables
    from crime_b200.gethi import GetHI, params_from_tables
    p = params_from_tables(load_tables(nu), n_grid=n, n_side=ns, seed=1001)
    out = {}
    with GetHI(p) as g:
        for _ in range(3):
            g.run(to_host=False)
        out["stage_ms"] = {k: round(v, 4) for k, v in g.stage_times().items()}
        ts = []
        for _ in range(5):
            g.zero_maps(); g.synchronize(); g.accumulate_maps(); g.synchronize()
            ts.append(g.stage_times()["maps"])
        out["maps_alone_ms"] = round(min(ts), 4)
        out["audit"] = {str(s): g.accumulate_audit(s) for s in (1.0, 0.5, 0.25)}
        m = g.run()
        out["maps_sum"] = float(np.asarray(m, dtype=np.float64).sum())
        np.save(os.environ["AB_OUT"], m)
    print("AB_RESULT " + json.dumps(out), flush=True)


def main():
    args = [a for a in sys.argv[1:] if "=" not in a]
    variants = [a for a in sys.argv[1:] if "=" in a]
    n, ns, nu = (int(a) for a in (args + ["512", "256", "64"][len(args):])[:3])
    import numpy as np
    import tempfile
    tmp = Path(tempfile.mkdtemp())
    base = None
    for v in [""] + variants:
        env = dict(os.environ, AB_CHILD="1", AB_OUT=str(tmp / f"m_{abs(hash(v))}.npy"))
        if v:
            k, val = v.split("=", 1)
            env[k] = val
        r = subprocess.run([sys.executable, __file__, str(n), str(ns), str(nu)], env=env, capture_output=True, text=True, timeout=900)
        line = [l for l in r.stdout.splitlines() if l.startswith("AB_RESULT ")]
        if not line:
            print(v or "default", "FAILED", r.stdout[-500:], r.stderr[-1500:])
            continue
        res = json.loads(line[0][len("AB_RESULT "):])
        m = np.load(env["AB_OUT"])
        if base is None:
            base = m
        same_set = bool(np.array_equal(m != 0, base != 0))
        nz = base != 0
        rel = float(np.abs(m[nz] / base[nz] - 1).max()) if same_set and nz.any() else float("nan")
        print(f"{v or 'default':28s} maps alone {res['maps_alone_ms']} ms | stages {res['stage_ms']} | lit pixels equal {same_set}, "
              f"max rel diff {rel:.2e} | audit wrong/unsure at 1, 1/2, 1/4: "
              + ", ".join(f"{a['wrong']}/{a['unsure']}" for a in res["audit"].values()), flush=True)


if __name__ == "__main__":
    if os.environ.get("AB_CHILD"):
        child(*(int(a) for a in sys.argv[1:4]))
    else:
        main()
