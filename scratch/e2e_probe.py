import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from bench import load_tables
from crime_b200.gethi import GetHI, params_from_tables
tables = load_tables(64)
params = params_from_tables(tables, n_grid=512, n_side=256, seed=1001)
g = GetHI(params)
stream = torch.cuda.ExternalStream(g.stream_handle())
def timed(fn, K, finish=True):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter(); e0.record(stream)
    for i in range(K): fn(i)
    t_host = time.perf_counter() - t0
    e1.record(stream)
    g.wait(); e2.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, e0.elapsed_time(e2) / K, 1e3 * t_host / K
def resident(i): g.run(to_host=False)
def full(i): g.set_params(params); g.run_async(i & 1)
def nosetp(i): g.run_async(i & 1)
def nocopy(i):
    g.set_params(params); g._check(g.lib.gh_cuda_run_async(g._ctx, None))
for _ in range(3): resident(0)
for name, fn in (("resident", resident), ("full", full), ("no set_params", nosetp), ("no copy", nocopy)):
    for K in (5, 20, 50):
        for i in range(2): fn(i)
        g.wait()
        print(f"{name:14s} K={K:3d} compute-stream ms/step {timed(fn, K)}", flush=True)
