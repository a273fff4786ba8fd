"""Worker for tests/test_gpu_multi.py: run under torchrun with one rank per GPU.

Every rank joins the slab-decomposed run (k-gen -> z FFT with the transpose fused in -> y/x FFT -> halo exchange ->
velocity -> variance all-reduce -> get_HI -> maps -> reduction by shell); rank 0 also runs the same problem alone on
its GPU and checks that the decomposed result is the same (crime_b200/selfcheck.py, which bench.py --gpus N also
runs before timing).

    multi_gpu_worker.py n_grid n_side [full|hash] [n_nu]
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from crime_b200 import params_from_tables  # noqa: E402
from crime_b200.selfcheck import decomposed_vs_single  # noqa: E402


def main():
    n_grid, n_side = int(sys.argv[1]), int(sys.argv[2])
    mode = sys.argv[3] if len(sys.argv) > 3 else "full"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    tables = dict(np.load(ROOT / "tests" / "golden" / "ref_tables_nu150.npz"))
    p = params_from_tables(tables, n_grid=n_grid, n_side=n_side, seed=31337)
    res = decomposed_vs_single(dist, p, rank, world, local, mode=mode, verbose=True)
    ok = True
    if rank == 0:
        ok = res["ok"]
        print("MULTI_GPU_RESULT " + json.dumps(res), flush=True)
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
