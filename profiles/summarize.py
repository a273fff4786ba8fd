"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

    python profiles/summarize.py r1 gpurun_out/launches_r1.csv gpurun_out/prof_r1_all.ncu-rep 512
"""
import collections
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("lts__t_sectors_op_red.sum", "L2 RED sectors"),
    ("lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "L2 atomic unit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
]


def short(name):
    return name.replace("<unnamed>::", "").replace("void ", "").split("(")[0]


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(short(r["Kernel Name"]), []).append((float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    # one step = the last occurrence block; count launches per step from the tail
    out = [f"# {tag}: kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none)", "",
           "Per-launch device times are cold-cache and serialised by ncu: compare SHARES, not absolutes.", "",
           "| kernel | launches captured | mean us | grid | block | share of captured time |", "|---|---|---|---|---|---|"]
    tot = sum(v[0] for k in per for v in per[k])
    for k, v in per.items():
        t = sum(x[0] for x in v)
        out.append(f"| `{k}` | {len(v)} | {t / len(v) / 1e3:.1f} | {v[-1][1]} | {v[-1][2]} | {100 * t / tot:.1f}% |")
    (HERE / f"{tag}_launches.md").write_text("\n".join(out) + "\n")
    (HERE / f"{tag}_launches.csv").write_text("".join(lines))


def full(tag, rep, n_grid):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    md = [f"# {tag}: ncu --set full --clock-control none, one launch per kernel ({n_grid}^3, 1 GPU)", ""]
    traffic = {}
    seen = set()
    for r in data:
        name = short(r[idx["Kernel Name"]])
        grid = r[idx["Grid Size"]]
        key = (name, grid)
        if key in seen:
            continue
        seen.add(key)
        md.append(f"## `{name}`  grid {grid} block {r[idx['Block Size']]}")
        md.append("")
        md.append("| metric | value |")
        md.append("|---|---|")
        for m, label in METRICS:
            if m in idx:
                md.append(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |")
        md.append("")
        try:
            def tobytes(m):
                v, u = float(r[idx[m]].replace(",", "")), units[idx[m]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            traffic.setdefault(name, tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"))
        except Exception:
            pass
    (HERE / f"{tag}_ncu_full.md").write_text("\n".join(md) + "\n")
    stage_of = {"kgen_kernel": "kgen", "radial_velocity_kernel": "vel", "get_HI_kernel": "get_HI", "accumulate_kernel": "maps",
                "sigma_partial_kernel": "sigma"}
    tf = HERE / "roofline_traffic.json"
    cur = json.loads(tf.read_text()) if tf.exists() else {}
    for k, v in traffic.items():
        base = k.split("<")[0]
        if base in stage_of:
            cur[f"{stage_of[base]}:{n_grid}"] = v
        cur[f"kernel:{base}:{n_grid}"] = v
    # fft stage = 2 fields x (2 strided + 1 rows)
    fs = [v for k, v in traffic.items() if k.startswith("fft_strided")]
    fr = [v for k, v in traffic.items() if k.startswith("fft_c2r_rows")]
    if fs and fr:
        cur[f"fft:{n_grid}"] = 2 * (2 * fs[0] + fr[0])
    tf.write_text(json.dumps(cur, indent=1) + "\n")


if __name__ == "__main__":
    tag, lcsv, rep, n = sys.argv[1:5]
    launches(tag, lcsv)
    full(tag, rep, int(n))
