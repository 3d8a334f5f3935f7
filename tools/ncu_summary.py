"""Condenses ncu output (read here, on the CPU box) into the small CSV / text files committed under profiles/.

  python tools/ncu_summary.py full <rep.ncu-rep> <out.csv>       selected metrics of every captured launch
  python tools/ncu_summary.py launches <launches.csv> <out.csv>  per-kernel totals + share of the launch list
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum$|gpu__dram_throughput\.avg\.pct|gpu__time_duration\.sum|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate\.pct|"
    r"lts__t_bytes\.sum$|launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_static|"
    r"occupancy_limit_\w+|waves_per_multiprocessor)$|sm__cycles_elapsed\.avg\.per_second|"
    r"sm__inst_executed_pipe_(alu|fma|fmaheavy|fp64|xu|lsu|uniform)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|smsp__inst_executed\.sum$|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed\.ratio|"
    r"sm__sass_inst_executed_op_(global|shared)_(ld|st)\.sum$)")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, launches = rows[0], rows[1], rows[2:]
    name_col = header.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
        w.writerow(["kernel", ""] + [r[name_col].split("(")[0][-60:] for r in launches])
        for c, h in enumerate(header):
            if KEEP.match(h):
                w.writerow([h, units[c]] + [r[c] for r in launches])
    print(f"{out}: {len(launches)} launches")


def launches(src, out):
    text = open(src).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", r["Kernel Name"])
        k = re.sub(r"^void |ocb::|at::native::|<unnamed>::", "", k)[:90]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_listed_gpu_time"])
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{t:.1f}", f"{t / n:.2f}", f"{t / total:.4f}"])
    print(f"{out}: {len(rows)} launches, {total / 1e3:.2f} ms listed")


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
