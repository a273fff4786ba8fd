import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from bench import load_tables
from crime_b200.gethi import GetHI, params_from_tables
tables = load_tables(64)
params = params_from_tables(tables, n_grid=512, n_side=256, seed=1001)
g = GetHI(params)
def resident(i): g.run(to_host=False)
def nosetp(i): g.run_async(i & 1)
for _ in range(3): resident(0)
g.wait(); print("resident", {k: round(v, 3) for k, v in g.stage_times().items()})
for K in (3, 10, 11):
    for i in range(K): nosetp(i)
    g.wait(); print("overlapped", K, {k: round(v, 3) for k, v in g.stage_times().items()})
