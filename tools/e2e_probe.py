"""Where does the end-to-end loop lose time against the resident one?  (run on a GPU box)
    python tools/e2e_probe.py [steps]
Times, with CUDA events on the library's stream: resident synchronous runs, resident asynchronous runs, asynchronous runs
with the maps copied to the host, the same plus gh_cuda_set_params every step (= bench.py's e2e loop)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from bench import load_tables  # noqa: E402
from crime_b200.gethi import GetHI, params_from_tables  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
p = params_from_tables(load_tables(64), n_grid=512, n_side=256, seed=1001)
with GetHI(p) as g:
    strm = torch.cuda.ExternalStream(g.stream_handle(), device=torch.device("cuda:0"))

    def timed(fn, finish=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(strm)
        for i in range(steps):
            fn(i)
        t_enq = time.perf_counter() - t0
        if finish:
            finish()
        e1.record(strm)
        g.synchronize()
        return e0.elapsed_time(e1) / steps, 1e3 * t_enq / steps

    for _ in range(3):
        g.run(to_host=False)
    print("resident, synchronous       ms/step %.3f  (host enqueue %.3f)" % timed(lambda i: g.run(to_host=False)))
    print("resident, asynchronous      ms/step %.3f  (host enqueue %.3f)" % timed(lambda i: g.run_async(None), finish=g.wait))
    for i in range(2):
        g.run_async(i & 1)
    g.wait()
    print("maps to host, asynchronous  ms/step %.3f  (host enqueue %.3f)" % timed(lambda i: g.run_async(i & 1), finish=g.wait))

    def full(i):
        g.set_params(p)
        g.run_async(i & 1)
    print("+ set_params every step     ms/step %.3f  (host enqueue %.3f)" % timed(full, finish=g.wait))
    t0 = time.perf_counter()
    for _ in range(20):
        g.set_params(p)
    g.synchronize()
    print("set_params alone: %.3f ms per call (host)" % (1e3 * (time.perf_counter() - t0) / 20))
