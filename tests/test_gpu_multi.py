"""Slab-decomposed run on >= 2 GPUs equals the single-GPU run (skipped on a 1-GPU box)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _ngpu():
    import torch
    return torch.cuda.device_count()


# bounds: GH_MAP_BOUNDS, forced plane ranges for the map accumulation (None = the cost model's; "off" = own slabs;
# "nccl" = the cost model's ranges with GH_NO_SPARSE_REDUCE=1: ncclReduceScatter instead of the default sparse map reduction
# over peer memory; "fused" / "onebuf" / "ownbuf": the other routes of the FFT transposes -- fused into the z pass as
# peer stores, copy engines with a single receive buffer, copy engines with a second buffer allocated for the purpose
# instead of the idle map stack).
# The forced cases make rank 0 pull planes from above, rank 1 from below, in more than one staging chunk, and
# leave one rank without any of its own planes.
@pytest.mark.parametrize("world,n_grid,n_side,bounds", [(2, 64, 32, None), (2, 64, 32, "57"), (2, 64, 32, "9"), (2, 64, 32, "off"),
                                                        (4, 128, 64, None), (4, 128, 64, "70,75,80"), (8, 128, 64, None),
                                                        (2, 64, 32, "nccl"), (4, 128, 64, "nccl"), (8, 128, 64, "nccl"),
                                                        (2, 64, 32, "fused"), (2, 64, 32, "onebuf"), (2, 64, 32, "ownbuf"), (4, 128, 64, "onebuf"),
                                                        (8, 128, 64, "fused"), (8, 128, 64, "ownbuf"),
                                                        (2, 64, 32, "nccl2"), (2, 64, 32, "nccl2onebuf"), (8, 128, 64, "nccl2"),
                                                        (2, 64, 32, "push"), (4, 128, 64, "push"), (8, 128, 64, "push")])
def test_slab_decomposition_matches_single_gpu(world, n_grid, n_side, bounds):
    import os
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    if bounds == "off":
        env["GH_NO_REBALANCE"] = "1"
    elif bounds == "nccl":  # the NCCL collective instead of the sparse map reduction over peer memory
        env["GH_NO_SPARSE_REDUCE"] = "1"
    elif bounds and bounds.startswith("nccl2"):  # the pipelined transposes as NCCL send/recv groups on a second communicator
        env["GH_TRANSPOSE"] = "nccl"
        if bounds.endswith("onebuf"):
            env["GH_ONE_RECV_BUFFER"] = "1"
    elif bounds == "push":  # the pipelined transposes as a store kernel over peer memory
        env["GH_TRANSPOSE"] = "push"
    elif bounds in ("fused", "onebuf", "ownbuf"):
        env[{"fused": "GH_FUSED_TRANSPOSE", "onebuf": "GH_ONE_RECV_BUFFER", "ownbuf": "GH_OWN_RECV_BUFFER"}[bounds]] = "1"
    elif bounds:
        env["GH_MAP_BOUNDS"] = bounds
    _run_worker(world, n_grid, n_side, "full", env, f"{world}gpu_{n_grid}_{bounds or 'model'}")


def _run_worker(world, n_grid, n_side, mode, env, tag):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + world), str(ROOT / "tests" / "multi_gpu_worker.py"), str(n_grid), str(n_side), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env)
    out = ROOT / "gpurun_out" / "multi_gpu"
    out.mkdir(parents=True, exist_ok=True)
    lines = [l for l in r.stdout.splitlines() if l.startswith(("MULTI_GPU_", "rank "))]
    (out / f"{tag.replace(',', '_')}.log").write_text("\n".join(lines) + "\n")
    assert "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world,n_grid,n_side", [(2, 1024, 512), (4, 1024, 512), (8, 2048, 1024)])
def test_benchmark_configurations_match_single_gpu(world, n_grid, n_side):
    """The BASELINE.json multi-GPU configurations themselves (150 shells): per-slab checksums of all five grids and
    the shell maps against the same run on one GPU (2048^3 needs 103 GiB on rank 0's device).  GH_TEST_LARGE=1."""
    import os
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    if not os.environ.get("GH_TEST_LARGE"):
        pytest.skip("large-grid multi-GPU check: GH_TEST_LARGE=1")
    _run_worker(world, n_grid, n_side, "hash", dict(os.environ), f"{world}gpu_{n_grid}_hash")


def test_c_host_fork_launcher_two_gpus(tmp_path):
    """GH_NGPUS=2 ./GetHI file: the C host forks one rank per GPU, passes the NCCL id through pipes, and every
    rank writes the shells it owns; the files equal a single-GPU run's."""
    import os
    import numpy as np
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, str(ROOT))
    from crime_b200 import host
    from oracle.binding import write_nutable, write_param_file
    write_nutable(tmp_path / "nu.txt", 10)
    outs = {}
    for tag, env in (("one", {}), ("two", {"GH_NGPUS": "2"})):
        write_param_file(tmp_path / f"{tag}.ini", n_grid=64, n_side=16, nutable=tmp_path / "nu.txt",
                         pk_file=ROOT / "data" / "Pk_synth.dat", prefix=tmp_path / tag, seed=12)
        r = subprocess.run([str(host.HOST_EXE), str(tmp_path / f"{tag}.ini")], capture_output=True, text=True, timeout=300,
                           env={**os.environ, **env})
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[tag] = [host.read_healpix_map(tmp_path / f"{tag}_{s + 1:03d}.fits")[0] for s in range(10)]
    for a, b in zip(outs["one"], outs["two"]):
        assert np.array_equal(a != 0, b != 0)
        nz = a != 0
        if nz.any():
            assert np.abs(b[nz] / a[nz] - 1).max() < 1e-5
