"""Condenses ncu output (read here, on the CPU box) into the small CSV / text files committed under profiles/.

  python tools/ncu_summary.py full <rep.ncu-rep> <out.csv>       selected metrics of every captured launch
  python tools/ncu_summary.py launches <launches.csv> <out.csv>  per-kernel totals + share of the launch list
  python tools/ncu_summary.py roofline <rep.ncu-rep> <out.json> [k1_variant] [comparisons per launch]
        the self-describing roofline record bench.py reads (profiles/k1_roofline.json): kernel name of the capture,
        variant index, hash of the kernel source and git revision it belongs to, DRAM bytes, pipe utilisation, and the
        per-comparison instruction mix counted in the SASS of that kernel's inner loop (tools/sass_loops.py)
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


def roofline(rep, out, variant="0", comparisons="100000000"):
    import hashlib
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import sass_loops
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, launches_ = rows[0], rows[2:]
    col = {h: c for c, h in enumerate(header)}
    k1 = [r for r in launches_ if "k1_top2_kernel" in r[col["Kernel Name"]]]
    if not k1:
        raise SystemExit("no k1_top2_kernel launch in " + rep)

    def avg(metric, scale=1.0):
        return sum(float(r[col[metric]].replace(",", "")) for r in k1) / len(k1) * scale

    unit = rows[1][col["dram__bytes_read.sum"]]
    to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    wunit = rows[1][col["dram__bytes_write.sum"]]
    kernel = k1[0][col["Kernel Name"]]
    args = re.search(r"<([^>]*)>", kernel).group(1).replace(" ", "").split(",")  # Q, F, IMADACC, COL, MINB, PFX, BF
    mangled = "k1_top2_kernelILi%sELi%sELb%sELb%sELi%sELb%sELb%sE" % tuple(args)
    lib = os.path.join(root, "opencalibration_b200", "libocb.so")
    name, dem, body = sass_loops.kernel_sass(lib, mangled)
    ins = sass_loops.parse(body)
    loops = []
    for addr, op, text in ins:
        if op.startswith("BRA"):
            m = re.search(r"(0x[0-9a-f]+)\s*$", text)
            if m and int(m.group(1), 16) <= addr:
                loops.append((int(m.group(1), 16), addr))

    def popc_density(l):
        inside = [i for i in ins if l[0] <= i[0] <= l[1]]
        k = sum(1 for i in inside if i[1].startswith("POPC"))
        return k / len(inside) if k >= 8 else 0.0
    lo, hi = max(loops, key=popc_density)
    inside = [i for i in ins if lo <= i[0] <= hi]
    q = int(args[0])
    per_trip = (4 if q <= 2 else 2) * q  # hamming_top2.cu: #pragma unroll(Q <= 2 ? 4 : 2) over candidates, Q queries per thread
    pipes = {}
    for _, op, _t in inside:
        pipes[sass_loops.pipe_of(op.split(".")[0])] = pipes.get(sass_loops.pipe_of(op.split(".")[0]), 0) + 1
    sha = lambda f: hashlib.sha256(open(os.path.join(root, "opencalibration_b200", "csrc", f), "rb").read()).hexdigest()
    git = subprocess.run(["git", "-C", root, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    rec = {
        "kernel": kernel.split("(")[0], "k1_variant": int(variant), "cross_check": args[3] in ("1", "true"),
        "source_sha256": {f: sha(f) for f in ("hamming_top2.cu",)}, "git_rev_of_capture_summary": git,
        "comparisons_per_launch": int(comparisons), "launches_averaged": len(k1), "report": os.path.basename(rep),
        "ncu": {"time_us": avg("gpu__time_duration.sum"),
                "dram_bytes": avg("dram__bytes_read.sum", to_bytes) + avg("dram__bytes_write.sum",
                                                                          {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[wunit]),
                "xu_pct_of_peak": avg("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                "alu_pct_of_peak": avg("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "fma_pct_of_peak": avg("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "issue_slots_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "sm_cycles_active": avg("sm__cycles_active.avg"), "sm_ghz": avg("sm__cycles_elapsed.avg.per_second"),
                "warp_instructions": avg("smsp__inst_executed.sum"),
                "registers_per_thread": avg("launch__registers_per_thread")},
        "sass": {"inner_loop": f"0x{lo:04x}..0x{hi:04x}", "instructions": len(inside), "comparisons_per_trip": per_trip,
                 "alu_ops_per_cmp": pipes.get("alu", 0) / per_trip, "xu_ops_per_cmp": pipes.get("xu", 0) / per_trip,
                 "fma_ops_per_cmp": pipes.get("fma", 0) / per_trip, "lsu_ops_per_cmp": pipes.get("lsu", 0) / per_trip},
    }
    # cross-check of the two sources: XU lane-operations per comparison as the counters saw them
    n = rec["ncu"]
    rec["ncu"]["xu_ops_per_cmp_from_counters"] = n["xu_pct_of_peak"] / 100 * 16 * 148 * n["sm_cycles_active"] / int(comparisons)
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))


def roofline_k1t(rep, out, comparisons="100000000"):
    """profiles/k1t_roofline.json: the tensor-core engine's three kernels (expansion, search, finish) of one step."""
    import hashlib
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, launches_ = rows[0], rows[1], rows[2:]
    col = {h: c for c, h in enumerate(header)}
    to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3}

    def val(r, metric):
        return float(r[col[metric]].replace(",", ""))

    kernels = {}
    for r in launches_:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        name = re.sub(r"^void |.*unnamed>::", "", name)
        k = kernels.setdefault(name, {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0, "rows": []})
        k["launches"] += 1
        k["time_us"] += val(r, "gpu__time_duration.sum") * to_us[units[col["gpu__time_duration.sum"]]]
        k["dram_bytes"] += (val(r, "dram__bytes_read.sum") * to_bytes[units[col["dram__bytes_read.sum"]]] +
                            val(r, "dram__bytes_write.sum") * to_bytes[units[col["dram__bytes_write.sum"]]])
        k["rows"].append(r)
    search = next(k for n, k in kernels.items() if re.search(r"k1t\d*_top2_kernel", n))
    steps = search["launches"]

    def avg(k, metric):
        return sum(val(r, metric) for r in k["rows"]) / len(k["rows"])
    sha = hashlib.sha256(open(os.path.join(root, "opencalibration_b200", "csrc", "hamming_tensor.cu"), "rb").read()).hexdigest()
    git = subprocess.run(["git", "-C", root, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    rec = {"kernels": {n: {"time_us": k["time_us"] / k["launches"], "dram_bytes": k["dram_bytes"] / k["launches"]}
                       for n, k in kernels.items()},
           "source_sha256": {"hamming_tensor.cu": sha}, "git_rev_of_capture_summary": git, "report": os.path.basename(rep),
           "comparisons_per_step": int(comparisons), "steps_averaged": steps,
           "ncu": {"dram_bytes_per_step": sum(k["dram_bytes"] for k in kernels.values()) / steps,
                   "search_time_us": search["time_us"] / steps,
                   "tensor_pipe_pct_active": avg(search, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   "alu_pct_of_peak": avg(search, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                   "issue_slots_pct": avg(search, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "l2_hit_pct": avg(search, "lts__t_sector_hit_rate.pct"),
                   "l2_to_sm_bytes": avg(search, "l1tex__m_xbar2l1tex_read_bytes.sum") *
                   to_bytes[units[col["l1tex__m_xbar2l1tex_read_bytes.sum"]]],
                   "tensor_core_shared_memory_wavefronts": avg(search, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum"),
                   "sm_cycles_active": avg(search, "sm__cycles_active.avg"),
                   "warp_instructions": avg(search, "smsp__inst_executed.sum"),
                   "registers_per_thread": avg(search, "launch__registers_per_thread"),
                   "sm_ghz": avg(search, "sm__cycles_elapsed.avg.per_second")}}
    rec["ncu"]["warp_instructions_per_32_comparisons"] = rec["ncu"]["warp_instructions"] / (int(comparisons) / 32)
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    {"full": full, "launches": launches, "roofline": roofline, "roofline_k1t": roofline_k1t}[sys.argv[1]](*sys.argv[2:])
