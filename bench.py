#!/usr/bin/env python
"""bench.py -- GetHI hot path, Mcells/s.

A step is one complete realisation through the hot path (k-space realisation, two c2r FFTs, radial
velocity, variance, lognormal/HI transform, HEALPix shell maps) of the configuration BASELINE.json names
for the GPU count:  1 GPU: 512^3, nside 256, 64 shells;  2/4 GPUs: 1024^3, nside 512, 150 shells;
8 GPUs: 2048^3, nside 1024, 150 shells (override with --grid/--nside/--shells).

  value : cells / device time with the run parameters already resident on the device and the maps left
          on the device (CUDA events on the library's stream, max over ranks) -- the device-resident figure
  e2e   : THE HEADLINE: the same through the reference-facing C-ABI with HOST buffers: parameter tables sent from
          host memory and this rank's finished maps copied back into pinned host memory inside the timed region
  roofline     : the dominant kernel timed alone, algorithmic bytes / duration vs measured HBM peak
  nvlink       : (N > 1) the FFT transposes (copy-engine peer copies by default): bytes shipped to peers / duration of the
                 transposition phase vs 900 GB/s
  parity       : (N > 1) a 128^3 slab-decomposed run against the same run on one GPU, done before timing
  strong_scaling : the 1024^3 / nside 512 / 150 shells problem timed at this N (fixed problem for the 1-2-4-8 curve)
  cpu_baseline : the reference's own CPU code (oracle/_ref) on the host cores (N = 1: the same 512^3 configuration)
  --impl reference : that CPU arm alone, same metric and config; each step is the 512^3 sample of the workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {1: (512, 256, 64), 2: (1024, 512, 150), 4: (1024, 512, 150), 8: (2048, 1024, 150)}
# algorithmic HBM bytes per cell of each stage (single-GPU FFT has no transpose pass)
# (the variance sums are produced inside the density FFT's last pass: no traffic of their own)
STAGE_BYTES_PER_CELL = {"kgen": 8.0, "fft": 32.0, "fft_multi": 48.0, "vel": 8.0, "sigma": 0.0, "get_HI": 16.0, "maps": 8.0}


def load_tables(n_nu: int) -> dict:
    f = ROOT / "tests" / "golden" / ("ref_tables_nu64.npz" if n_nu == 64 else "ref_tables_nu150.npz")
    t = dict(np.load(f))
    if int(t["n_nu"]) != n_nu:  # other shell counts: n+1 uniform edges over the same band, %.6f like data/nuTable.txt
        edges = np.array([float(f"{e:.6f}") for e in np.linspace(355.0, 945.0, n_nu + 1)])
        t["nu0_arr"], t["nuf_arr"], t["n_nu"] = edges[:-1].copy(), edges[1:].copy(), np.asarray(n_nu)
    return t


def t_total_c_host(n_grid: int, n_side: int, n_nu: int, n_gpus: int = 1) -> dict:
    """T_total as the reference's own total timer brackets it (main_gh.c:42,69) -- `./GetHI file` from
    param read through cosmology tables, device bring-up and the hot path to the last FITS map on disk (tmpfs),
    with this repository's C host (host/GetHI): the maps are written while they are still being downloaded."""
    import re
    import shutil
    import subprocess
    import tempfile
    exe = ROOT / "host" / "GetHI"
    if not exe.exists():
        return {"error": "host/GetHI not built"}
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    d = Path(tempfile.mkdtemp(prefix="gethi_ttotal_", dir=base))
    try:
        edges = np.linspace(355.0, 945.0, n_nu + 1)
        (d / "nu.txt").write_text("".join(f"{e:.6f}\n" for e in edges))
        (d / "p.ini").write_text(
            f"prefix_out= {d}/map\npk_filename= {ROOT}/data/Pk_synth.dat\nomega_M= 0.3\nomega_L= 0.7\nomega_B= 0.049\nh= 0.67\n"
            f"w= -1.0\nns= 0.96\nsigma_8= 0.8\nr_smooth= 2.0\nfrequencies_filename= {d}/nu.txt\nn_side= {n_side}\n"
            f"n_grid= {n_grid}\nseed= 1001\ndo_psources= 0\n")
        t0 = time.perf_counter()
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "GH_RANK", "GH_NRANKS"):  # the C host launches its own ranks
            env.pop(k, None)
        env["GH_HOST_TIMING"] = "1"
        if n_gpus > 1:
            env["GH_NGPUS"] = str(n_gpus)
        r = subprocess.run([str(exe), str(d / "p.ini")], capture_output=True, text=True, timeout=600, env=env)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": (r.stdout + r.stderr)[-300:]}
        m = re.search(r"Total time ellapsed ([0-9.]+) ms", r.stdout)
        written = sum(f.stat().st_size for f in d.glob("map_*.fits"))
        inside = float(m.group(1)) / 1e3 if m else None
        phases = {m_.group(1): float(m_.group(2)) for m_ in re.finditer(r"\[gh_host\] (.+?) ([0-9.]+) ms", r.stderr)}
        return {"process_wall_s": round(wall, 3), "reference_total_timer_s": inside, "fits_bytes_written": written, "n_gpus": n_gpus,
                "host_phase_ms": phases,
                "mcells_per_s": (float(n_grid) ** 3 / inside / 1e6) if inside else None,
                "what": "host/GetHI param file -> FITS maps on tmpfs: param read, cosmology tables, device bring-up, hot path, "
                        f"{n_nu} maps written by the streaming writer"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in Path(self.path).read_text().splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def bind_to_gpu_numa_node(device_index: int) -> None:
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, so that the page-locked map
    buffers are first-touched on the NUMA node next to the GPU's PCIe root (host-side plumbing only)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def measured_peaks() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(n_grid: int, n_side: int, n_nu: int, steps: int, warmup: int) -> dict:
    """Time the reference's own create_d_and_vr_fields + get_HI + mk_T_maps (oracle/_ref, all host
    threads); falls back to the oracle port when the prebuilt reference is absent."""
    from oracle.binding import Oracle, Reference, write_nutable, write_param_file
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every rank unless the user set it; the CPU arm runs on rank 0 alone and
    # is meant to use every host thread it can, so that default is overridden (an explicit GH_CPU_THREADS wins)
    cores = int(os.environ.get("GH_CPU_THREADS", "0")) or avail
    if "LOCAL_RANK" not in os.environ and os.environ.get("OMP_NUM_THREADS"):
        cores = int(os.environ["OMP_NUM_THREADS"])
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:  # libgomp may already be initialised in this process: set the thread count through its API as well
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except OSError:
        pass
    cells = float(n_grid) ** 3
    times, stage = [], []
    if Reference.available():
        kind = "reference"
        ref = Reference()
        tmp = tempfile.mkdtemp()
        write_nutable(f"{tmp}/nu.txt", n_nu)
        write_param_file(f"{tmp}/p.ini", n_grid=n_grid, n_side=n_side, nutable=f"{tmp}/nu.txt",
                         pk_file=str(ROOT / "data" / "Pk_synth.dat"), prefix=f"{tmp}/out")
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)  # the reference prints its banner and timers on stdout
        try:
            par = ref.read_run_params(f"{tmp}/p.ini")
            maps = ref.grid(par, "maps_HI", (n_nu, 12 * n_side * n_side))   # allocated once by the reference's reader
            for i in range(warmup + steps):
                maps[:] = 0                                                 # outside the timed region
                t0 = time.perf_counter()
                ref.lib.ref_create_d_and_vr_fields(par)
                t1 = time.perf_counter()
                ref.lib.ref_get_HI(par)
                t2 = time.perf_counter()
                ref.lib.ref_mk_T_maps(par)
                t3 = time.perf_counter()
                if i >= warmup:
                    times.append(t3 - t0)
                    stage.append((t1 - t0, t2 - t1, t3 - t2))
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
    else:
        kind = "port"
        from crime_b200 import params_from_tables
        orc = Oracle()
        p = params_from_tables(load_tables(n_nu), n_grid=n_grid, n_side=n_side)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            orc.run(p)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    st = np.mean(np.asarray(stage), axis=0) if stage else None
    return {"value": cells / t / 1e6, "unit": "Mcells/s", "cores": cores, "kind": kind, "seconds_per_step": t,
            "stage_seconds": None if st is None else {"create_d_and_vr_fields": float(st[0]), "get_HI": float(st[1]), "mk_T_maps": float(st[2])},
            "sample": f"{n_grid}^3 grid, nside {n_side}, {n_nu} shells, {len(times)} step(s); FFT stage runs the oracle's "
                      f"own C FFT (FFTW is not installed), every other stage is the reference's code"}


METRIC = "GetHI Mcells/s end-to-end"
PK_NOTE = ("data/Pk_synth.dat: a synthetic CAMB-like linear P(k) (data/make_synthetic_pk.py) standing in for the reference's "
           "data/Pk_CAMB_test.dat, which is not redistributed here; both arms read the same file")


def make_config(n_grid: int, n_side: int, n_nu: int, world: int) -> dict:
    """The `config` object: identical in both arms so that the driver compares like with like."""
    cells = float(n_grid) ** 3
    return {"workload": f"GetHI {n_grid}^3 grid, nside={n_side}, {n_nu} shells", "n_grid": n_grid, "n_side": n_side,
            "n_nu": n_nu, "slabs": world, "cells_per_gpu": cells / world, "pk_file": PK_NOTE,
            "l2": "inputs larger than L2: three %.2f GiB grids per GPU are swept every step" % (cells / world * 4 * (1 + 2.0 / n_grid) / 2**30)}


def cpu_sample_of(n_grid: int, n_side: int, n_nu: int, cpu_grid: int = 0):
    """The bounded sample the CPU arm times: the workload itself up to 512^3; larger grids are sampled at 512^3 with
    n_side scaled by the same factor (same cells per pixel, same sub-particles per pixel) and the same shells."""
    sg = cpu_grid or min(n_grid, 512)
    ns = max(1, (n_side * sg) // n_grid) if sg < n_grid else n_side
    return sg, ns, n_nu


def run_reference_arm(args, workload):
    n_grid, n_side, n_nu = workload
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sg, ns, nn = cpu_sample_of(n_grid, n_side, n_nu, args.cpu_grid)
    r = cpu_reference_run(sg, ns, nn, args.steps, args.warmup)
    whole = (sg, ns, nn) == (n_grid, n_side, n_nu)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Mcells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 math, f32 storage",
        "data": "synthetic",
        "config": make_config(n_grid, n_side, n_nu, args.gpus),
        "sample": {"n_grid": sg, "n_side": ns, "n_nu": nn, "whole_workload": whole,
                   "what": "the workload itself" if whole else
                   f"a {sg}^3 sample of the workload: n_side scaled with the grid ({ns}), same shells; Mcells/s of the sample"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_seconds": r.get("stage_seconds"),
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--nside", type=int, default=0)
    ap.add_argument("--shells", type=int, default=0)
    ap.add_argument("--cpu-grid", type=int, default=0, help="grid of the bounded CPU sample (default min(grid,512))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-t-total", action="store_true", help="skip the ./GetHI param-file-to-FITS run")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arm (memory-capacity stress configs)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the 128^3 decomposed-vs-single-GPU check before timing")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed 1024^3 problem timed for the strong-scaling curve")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    wl = list(WORKLOADS.get(args.gpus, WORKLOADS[1]))
    if args.grid: wl[0] = args.grid
    if args.nside: wl[1] = args.nside
    if args.shells: wl[2] = args.shells
    n_grid, n_side, n_nu = wl
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    from crime_b200 import GetHI, params_from_tables
    from crime_b200 import abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GetHI hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(local_rank)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        buf = torch.zeros(abi.GH_CUDA_UNIQUE_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            import ctypes as C
            raw = C.create_string_buffer(abi.GH_CUDA_UNIQUE_ID_BYTES)
            lib = abi.load_library()
            if lib.gh_cuda_get_unique_id(raw):
                raise SystemExit(lib.gh_cuda_last_error().decode())
            buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        buf = buf.to(dev)
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    # ---- parity before timing (N > 1): a 128^3 slab-decomposed run against the same run on one GPU --------------
    parity = None
    if world > 1 and not args.no_parity:
        from crime_b200.selfcheck import decomposed_vs_single
        pp = params_from_tables(load_tables(150), n_grid=128, n_side=64, seed=31337)
        try:
            parity = decomposed_vs_single(dist, pp, rank, world, local_rank, mode="full")
        except Exception as exc:  # a failed check must show up in the line, not kill the measurement
            parity = {"ok": False, "error": str(exc)[:300]} if rank == 0 else None
        barrier()

    if world > 1:
        os.environ.setdefault("GH_TIME_FFT_PASSES", "1")  # events around the z passes: the NVLink figure below
    tables = load_tables(n_nu)
    params = params_from_tables(tables, n_grid=n_grid, n_side=n_side, seed=1001)
    g = GetHI(params, rank=rank, nranks=world, unique_id=uid, device=local_rank)
    stream = torch.cuda.ExternalStream(g.stream_handle(), device=dev)
    cells = float(n_grid) ** 3

    def timed(gh, strm, fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(strm)
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        e1.record(strm)
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), wall

    def step_resident(i=0):
        g.run(to_host=False)

    def step_e2e(i=0):
        # host -> device: the run's parameter block and tables; device -> host: this rank's finished maps into
        # one of two pinned buffers.  Nothing waits for the copy until the end (gh_cuda_run_async /
        # gh_cuda_wait): the copy of realisation i overlaps the computation of realisation i+1.
        g.set_params(params)
        g.run_async(i & 1)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = g.kernel_launches()
    ms_res, _ = timed(g, stream, step_resident, args.steps)
    launches = g.kernel_launches() - l0
    if args.no_e2e:
        ms_e2e, wall_e2e = float("nan"), float("nan")
    else:
        for i in range(2):
            step_e2e(i)
        g.wait()
        ms_e2e, wall_e2e = timed(g, stream, step_e2e, args.steps, finish=g.wait)
        g.run(to_host=True)  # one synchronous realisation so that the per-stage timers below include an unoverlapped copy
    clocks = sampler.stop()
    stage_ms = g.stage_times()
    z_pass = g.fft_pass_times() if world > 1 else None

    # dominant kernel alone: every stage once more, each bracketed by its own events; on several ranks a barrier in
    # front of each stage so that a stage's time is not somebody else's lateness
    solo = {}
    for name, call in (("kgen", g.generate_k), ("fft", g.fft_fields), ("vel", g.radial_velocity), ("sigma", g.sigma_dens),
                       ("get_HI", g.get_HI), ("zero", g.zero_maps), ("maps", g.accumulate_maps)):
        g.synchronize()
        barrier()
        call()
        g.synchronize()
        if name != "zero":
            solo[name] = g.stage_times()[name]
    if world > 1:  # the slowest rank's time of each stage
        tt = torch.tensor([solo[k] for k in sorted(solo)], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        solo = dict(zip(sorted(solo), [float(v) for v in tt.tolist()]))
    nz_cells = cells / world
    bytes_per_cell = dict(STAGE_BYTES_PER_CELL)
    if world > 1:
        bytes_per_cell["fft"] = bytes_per_cell["fft_multi"]
    cand = {k: solo[k] for k in ("kgen", "fft", "vel", "sigma", "get_HI", "maps") if solo.get(k, 0) > 0}
    top = max(cand, key=cand.get)
    peak, peak_src = measured_peaks()
    achieved = bytes_per_cell[top] * nz_cells / (cand[top] * 1e-3) / 1e9
    fft_strided = "fft_strided_tma_kernel" if (n_grid <= 512 and not os.environ.get("GH_FFT_NO_TMA")) or os.environ.get("GH_FFT_TMA") else "fft_strided_kernel"
    kernel_names = {"kgen": "kgen_kernel", "fft": f"{fft_strided} x4 + fft_c2r_rows_kernel x2", "vel": "radial_velocity_kernel",
                    "sigma": "sigma_partial_kernel", "get_HI": "get_HI_kernel", "maps": "accumulate_kernel"}
    traffic = None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(f"{top}:{n_grid}" if world == 1 else f"{top}:{n_grid}:{world}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": kernel_names[top], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_cell": bytes_per_cell[top], "ms": cand[top],
                "stage_ms_alone": {k: round(v, 4) for k, v in solo.items()},
                "stage_frac_of_hbm_peak": {k: round(bytes_per_cell[k] * nz_cells / (v * 1e-3) / 1e9 / peak, 4) for k, v in cand.items()},
                "timed": "each stage alone after a stream synchronise" + (" and a cross-rank barrier; slowest rank" if world > 1 else "")}
    if top == "maps":
        n_sub = 10.0 * nz_cells
        roofline["note"] = ("accumulate_kernel is bound by instruction issue and L2 atomic throughput, not HBM "
                            "(SURVEY 8d): also reporting sub-particles/s")
        roofline["subparticles_per_s"] = n_sub / (cand[top] * 1e-3)

    per_rank = None
    if world > 1:
        # every rank's own stage timers of the last pipelined step: shows load imbalance between slabs
        mine = [stage_ms.get(k, 0.0) for k in abi.STAGE_NAMES]
        tt = torch.tensor(mine, device=dev, dtype=torch.float64)
        allt = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allt, tt)
        per_rank = [{k: round(float(v), 3) for k, v in zip(abi.STAGE_NAMES, a.tolist())} for a in allt]
    nvlink = None
    transpose_route = ("fused into the z pass" if os.environ.get("GH_FUSED_TRANSPOSE") or os.environ.get("GH_TRANSPOSE") == "fused"
                       else "nccl send/recv, second communicator" if os.environ.get("GH_TRANSPOSE") == "nccl"
                       else "store kernel over peer memory" if os.environ.get("GH_TRANSPOSE") == "push" else "copy engines")
    if world > 1:
        # the transposition phase of each field: every rank ships (P-1)/P of its slab to its peers -- copy-engine peer
        # copies by default, the fused z pass with GH_FUSED_TRANSPOSE=1.  Slowest rank's time, so the figure is the
        # all-to-all's, not one link's.
        zt = torch.tensor(list(z_pass), device=dev, dtype=torch.float64)
        dist.all_reduce(zt, op=dist.ReduceOp.MAX)
        z_ms = [float(v) for v in zt.tolist()]
        sent = (world - 1) / world * (cells / world) * (1 + 2.0 / n_grid) * 4.0   # bytes per rank per field
        if min(z_ms) > 0:
            ach = 2 * sent / (sum(z_ms) * 1e-3) / 1e9
            nvlink = {"achieved": ach, "peak": 900.0, "unit": "GB/s per GPU per direction", "frac": ach / 900.0,
                      "z_pass_ms": z_ms, "bytes_sent_per_rank_per_field": sent,
                      "route": transpose_route,
                      "what": ("copy-engine peer copies of the in-place z pass's output (one contiguous block per peer), first "
                               "copy start to last copy end per field, overlapped with the other field's FFT passes; last e2e step; "
                               "peak = NVLink 5 nominal" if transpose_route == "copy engines" else
                               "transposition phase as timed by the library (fused: the z pass itself, which also reads and "
                               "transforms the slab, so a lower bound on the link rate), last e2e step; peak = NVLink 5 nominal")}
        else:
            nvlink = {"achieved": None, "peak": 900.0, "unit": "GB/s per GPU per direction", "frac": None, "z_pass_ms": z_ms,
                      "what": "z-pass timers unavailable"}
    n_here = g.n_shells_here
    table_bytes = sum(np.asarray(v).nbytes for k, v in tables.items() if k in abi.TABLE_FIELDS)
    maps_bytes = n_here * g.npix * 4
    if world > 1:
        mb = torch.tensor([float(maps_bytes), float(table_bytes)], device=dev, dtype=torch.float64)
        dist.all_reduce(mb)
        maps_bytes, table_bytes = int(mb[0].item()), int(mb[1].item())
    g.end_fftw()

    # ---- fixed problem for the strong-scaling curve: 1024^3 / nside 512 / 150 shells at this N ------------------------
    strong = None
    SS = (1024, 512, 150)
    if not args.no_strong:
        if (n_grid, n_side, n_nu) == SS:
            strong = {"workload": "GetHI 1024^3 grid, nside=512, 150 shells", "ms_per_step": ms_res / args.steps,
                      "value": cells * args.steps / (ms_res * 1e-3) / 1e6, "e2e_value": cells * args.steps / (ms_e2e * 1e-3) / 1e6,
                      "steps": args.steps, "note": "this N's main workload"}
        else:
            try:
                uid2 = None
                if world > 1:
                    from crime_b200.selfcheck import make_unique_id
                    uid2 = make_unique_id(dist, rank)
                ps = params_from_tables(load_tables(SS[2]), n_grid=SS[0], n_side=SS[1], seed=1001)
                gs = GetHI(ps, rank=rank, nranks=world, unique_id=uid2, device=local_rank)
                ss_stream = torch.cuda.ExternalStream(gs.stream_handle(), device=dev)
                k = max(3, min(args.steps, 5))
                for _ in range(3):
                    gs.run(to_host=False)
                ms_s, _ = timed(gs, ss_stream, lambda i: gs.run(to_host=False), k)

                def ss_e2e(i):
                    gs.set_params(ps)
                    gs.run_async(i & 1)
                for i in range(2):
                    ss_e2e(i)
                gs.wait()
                ms_se, _ = timed(gs, ss_stream, ss_e2e, k, finish=gs.wait)
                gs.end_fftw()
                c3 = float(SS[0]) ** 3
                strong = {"workload": "GetHI 1024^3 grid, nside=512, 150 shells", "ms_per_step": ms_s / k, "value": c3 * k / (ms_s * 1e-3) / 1e6,
                          "e2e_value": c3 * k / (ms_se * 1e-3) / 1e6, "steps": k, "note": "timed after the main workload, same process"}
            except Exception as exc:
                strong = {"error": str(exc)[:300]}

    if world > 1:
        # the other ranks leave now (releasing their GPUs): what follows on rank 0 -- the C host's own multi-GPU run for
        # T_total -- must not share the devices with ranks spinning in a barrier
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    if rank == 0:
        line = {
            "metric": METRIC, "value": cells * args.steps / (ms_res * 1e-3) / 1e6, "unit": "Mcells/s",
            "value_definition": "device-resident: cells / CUDA-event time, parameters on the device, maps left there; "
                                "`e2e` is the end-to-end figure (host tables in, host maps out, every step) and the headline",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 fields, f64 index arithmetic",
            "data": "synthetic",
            "config": make_config(n_grid, n_side, n_nu, world),
            "e2e": ({"value": None, "unit": "Mcells/s", "h2d_bytes_per_step": table_bytes, "d2h_bytes_per_step": maps_bytes,
                     "note": "--no-e2e"} if args.no_e2e else
                    {"value": cells * args.steps / (ms_e2e * 1e-3) / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": table_bytes,
                     "d2h_bytes_per_step": maps_bytes, "ms_per_step": ms_e2e / args.steps,
                     "wall_ms_per_step": 1e3 * wall_e2e / args.steps}),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "stage_ms_last_e2e_step": {k: round(v, 4) for k, v in stage_ms.items()},
        }
        if per_rank is not None:
            line["stage_ms_by_rank"] = per_rank
        if nvlink is not None:
            line["nvlink"] = nvlink
        if parity is not None:
            line["parity"] = parity
        if strong is not None:
            line["strong_scaling"] = strong
        if not args.no_cpu_baseline and world == 1:
            sg, ns, nn = cpu_sample_of(n_grid, n_side, n_nu, args.cpu_grid)
            try:
                cb = cpu_reference_run(sg, ns, nn, 2, 1)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as exc:  # the CPU arm must never take the GPU line down with it
                line["cpu_baseline"] = {"value": None, "unit": "Mcells/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {exc}"}
        if not args.no_t_total:
            try:
                line["t_total"] = t_total_c_host(n_grid, n_side, n_nu, world)
            except Exception as exc:
                line["t_total"] = {"error": str(exc)}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
