#!/usr/bin/env python
"""Turns ncu captures brought back in gpurun_out/ into the committed, human-readable summaries under profiles/.

    python profiles/summarize.py full   gpurun_out/<name>.ncu-rep  profiles/<name>_summary.md   [--traffic-key "kernel=bench name" ...]
    python profiles/summarize.py launch gpurun_out/<name>.csv      profiles/<name>_launches.md

`full` reads `ncu -i <rep> --page raw --csv`; `launch` reads the `--metrics gpu__time_duration.sum` launch list.
With --traffic-key it also records dram read+write bytes per launch in profiles/traffic.json (bench.py's roofline.traffic).
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM read bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy achieved)"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__waves_per_multiprocessor", "waves / SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def full(rep, out, traffic_keys):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full summary: %s" % os.path.basename(rep), "",
             "Captured on a B200 through gpurun with `--clock-control none`; per-launch values (cold cache, serialised replay).", ""]
    traffic = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        lines.append("## `%s`  grid %s block %s" % (name[:90], r[idx["Grid Size"]], r[idx["Block Size"]]))
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for key, label in METRICS:
            if key in idx and r[idx[key]] != "":
                lines.append("| %s (`%s`) | %s %s |" % (label, key, r[idx[key]], units[idx[key]]))
        stalls = sorted(((float(r[i]), h) for h, i in idx.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and
                         not h.endswith("_not_issued") and r[i] not in ("", "0")), reverse=True)[:6]
        if stalls:
            lines.append("| top stall reasons (pc samples) | %s |" % ", ".join("%s %d" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), v) for v, h in stalls))
        lines.append("")
        for spec in traffic_keys:
            pattern, bench_name = spec.split("=", 1)
            if pattern in name + " " + r[idx["Grid Size"]] and bench_name not in traffic:
                traffic[bench_name] = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                    to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    open(out, "w").write("\n".join(lines) + "\n")
    if traffic:
        path = os.path.join(os.path.dirname(os.path.abspath(out)), "traffic.json")
        existing = json.load(open(path)) if os.path.exists(path) else {}
        existing.update(traffic)
        json.dump(existing, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", out, traffic)


def launch(csv_path, out):
    rows = list(csv.reader(open(csv_path)))
    start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[start]
    ki, gi, bi, mi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size"), hdr.index("Metric Name"), hdr.index("Metric Value")
    groups = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) > vi and r[mi] == "gpu__time_duration.sum":
            groups.setdefault((r[ki][:70], r[gi], r[bi]), []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in groups.values())
    lines = ["# ncu launch list: %s" % os.path.basename(csv_path), "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` over a short bench.py run; times are cold-cache and",
             "serialised, so compare SHARES with the live CUDA-event numbers in the bench JSON, not absolutes.", "",
             "| kernel | grid | block | launches | avg us | share of listed time |", "|---|---|---|---|---|---|"]
    for (name, grid, block), v in groups.items():
        lines.append("| `%s` | %s | %s | %d | %.1f | %.1f%% |" % (name, grid, block, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / total))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    keys = [a for a in sys.argv[4:] if not a.startswith("--")]
    if mode == "full":
        full(src, dst, keys)
    else:
        launch(src, dst)
