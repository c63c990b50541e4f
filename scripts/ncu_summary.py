#!/usr/bin/env python
"""Condenses ncu output into the small text/JSON files kept under profiles/.

    python scripts/ncu_summary.py report  <file.ncu-rep> <out.txt> [workload-key]   # one `--set full` capture
    python scripts/ncu_summary.py launches <launches.csv> <out.txt>                  # a gpu__time_duration launch list

`report` also merges dram bytes per launch into profiles/traffic.json under the workload key (read by bench.py's
roofline.traffic).  Runs here, without a GPU (ncu -i only imports the report).
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"l1tex__data_pipe_lsu_wavefronts(_mem_shared(_op_(ld|st))?)?\.(sum|avg)(\.pct_of_peak_sustained_elapsed)?|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st))?\.sum|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_sector_hit_rate\.pct|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__cycles_elapsed\.avg|sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__inst_executed\.sum|smsp__thread_inst_executed_per_inst_executed\.ratio|smsp__warps_eligible\.avg\.per_cycle_active|"
    r"smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio|sm__pipe_fp64_cycles_active.*pct_of_peak_sustained_elapsed|"
    r".*sm__pipe_fp64_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed|sm__inst_executed_pipe_(fp64|lsu|alu|fma|fmaheavy|uniform|xu|cbu|adu)\.sum|"
    r"smsp__sass_inst_executed_op_(shared_ld|shared_st|global_ld|global_st|tma_ld|tma_st|local_ld|local_st)\.sum|"
    r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit_[a-z_]+|waves_per_multiprocessor)|"
    r"smsp__cycles_active\.avg)$")


def report(path, out, key=None):
    if path.endswith(".csv"):      # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`): the report itself stayed there
        with open(path) as f:
            raw = f.read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, unit = rows[0], rows[1]
    lines = []
    traffic = None
    for row in rows[2:]:
        d = OrderedDict((h, (u, v)) for h, u, v in zip(head, unit, row))
        name = d.get("Kernel Name", ("", "?"))[1]
        lines.append(f"kernel: {name}")
        lines.append(f"grid {d.get('Grid Size', ('', '?'))[1]}  block {d.get('Block Size', ('', '?'))[1]}")
        for h, (u, v) in d.items():
            short = h.split(".TriageCompute.")[-1] if ".TriageCompute." in h else h
            if KEEP.match(short) and v not in ("", "0", "0.000000"):
                lines.append(f"  {short:<95s} {v} {u}")
        def num(k):
            u, v = d[k]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
            return float(v) * scale
        try:
            traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
            lines.append(f"  dram traffic per launch (read+write): {traffic:.0f} bytes")
        except Exception:
            pass
        lines.append("")
    with open(out, "w") as f:
        f.write(f"# condensed from {os.path.basename(path)} (ncu --set full --clock-control none --import-source on)\n")
        f.write("\n".join(lines))
    if key and traffic is not None:
        tj = os.path.join(ROOT, "profiles", "traffic.json")
        cur = {}
        if os.path.exists(tj):
            with open(tj) as f:
                cur = json.load(f)
        cur[key] = traffic
        with open(tj, "w") as f:
            json.dump(cur, f, indent=1, sort_keys=True)
    print("\n".join(lines))


def launches(path, out):
    with open(path) as f:
        text = f.read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        k = re.sub(r"\(.*", "", r["Kernel Name"])[:160]
        n, t, mx = per.get(k, (0, 0.0, 0.0))
        per[k] = (n + 1, t + ns, max(mx, ns))
    total = sum(t for _, t, _ in per.values())
    lines = [f"# condensed from {os.path.basename(path)}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)",
             f"# {sum(n for n, _, _ in per.values())} launches, {total / 1e6:.3f} ms of device time", "",
             f"{'share':>7s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'max_us':>9s}  kernel"]
    for k, (n, t, mx) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{100 * t / total:6.2f}% {n:8d} {t / 1e3:10.1f} {t / 1e3 / n:9.2f} {mx / 1e3:9.2f}  {k}")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "report":
        report(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        launches(sys.argv[2], sys.argv[3])
