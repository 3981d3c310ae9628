#!/usr/bin/env python
"""Turn the ncu outputs of a gpurun call into the small CSVs kept under profiles/.

  python tools/ncu_summarize.py launches <launches.csv> <out_by_kernel.csv>
      per-kernel aggregation of an `ncu --metrics gpu__time_duration.sum --csv` launch list
  python tools/ncu_summarize.py full <report.ncu-rep>[,<report2>...] <out_summary.csv> [kernels,to,skip,in,the,first]
      one row per captured launch of `ncu --set full` reports (duration, grid, registers, shared memory, achieved
      occupancy, DRAM bytes read/written, DRAM and SM throughput, L2 bytes, instructions, shared-memory bank conflicts)
  python tools/ncu_summarize.py hotspots <report.ncu-rep> <out.txt> <kernel,regexes>
      per kernel, the SASS instructions with the most warp-stall samples and their stall reasons
  python tools/ncu_summarize.py roofline <full_summary.csv> <out_roofline.csv>
      derived: DRAM GB/s per launch = (read + written bytes) / duration and its fraction of the HBM peak
"""
import collections
import csv
import re
import subprocess
import sys

FULL_COLS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], f"{a[1]:.1f}", f"{a[1] / a[0]:.1f}", f"{a[1] / total:.4f}"])
    print(f"{len(agg)} kernels, {total / 1000:.3f} ms of kernel time -> {dst}")


_TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}  # -> us
_BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}  # -> Mbyte
_SMEM = {"byte/block": 1e-3, "Kbyte/block": 1.0, "Mbyte/block": 1e3}  # -> Kbyte/block


def _normalise(value, unit):
    """ncu picks a unit per column and report: bring times to us, byte counts to Mbyte, shared memory to Kbyte/block."""
    try:
        v = float(value.replace(",", ""))
    except ValueError:
        return value, unit
    if unit in _TIME:
        return f"{v * _TIME[unit]:.3f}", "us"
    if unit in _BYTES:
        return f"{v * _BYTES[unit]:.6f}", "Mbyte"
    if unit in _SMEM:
        return f"{v * _SMEM[unit]:.3f}", "Kbyte/block"
    return value, unit


def full(reps, dst, skip=()):
    """One row per captured launch of one or several reports (comma separated), units normalised; kernels whose name starts
    with an entry of `skip` are dropped (stale rows of an older capture that a newer report replaces)."""
    out_rows, out_units, cols = [], None, None
    for rep in reps.split(","):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = [hdr.index(c) for c in FULL_COLS if c in hdr]
        cols = [hdr[i] for i in idx]
        for d in data:
            name = d[idx[0]].replace("void ", "")
            if any(name.startswith(k) for k in skip):
                continue
            vals, us = [], []
            for i in idx:
                v, u = _normalise(d[i][:70], units[i])
                vals.append(v)
                us.append(u)
            out_units = us
            out_rows.append(vals)
        skip = ()  # only the first report is filtered
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow(out_units)
        w.writerows(out_rows)
    print(f"{len(out_rows)} launches -> {dst}")


def hotspots(rep, dst, kernels):
    """Per kernel: the instructions with the most warp-stall samples (SASS, needs -lineinfo / --import-source for the source
    page), with their dominant stall reasons and their share of the kernel's samples."""
    lines_out = []
    for kern in kernels:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern, "-c", "1"], capture_output=True, text=True).stdout
        lines = out.splitlines()
        start = next((i for i, l in enumerate(lines) if l.startswith('"Address"')), None)
        if start is None:
            continue
        name = lines[0].split('","')[1][:90] if lines and '","' in lines[0] else kern
        r = csv.reader(lines[start:])
        hdr = next(r)
        ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        stalls = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        rows = []
        for row in r:
            if len(row) != len(hdr) or row[0] == "Address":
                break
            rows.append(row)
        total = sum(int(x[isamp]) for x in rows) or 1
        by_reason = collections.Counter()
        for row in rows:
            for j, h in stalls:
                by_reason[h] += int(row[j])
        lines_out.append(f"== {name}  ({len(rows)} SASS instructions, {total} stall samples)")
        lines_out.append("   stall reasons: " + ", ".join(f"{h[6:]} {100.0 * c / max(1, sum(by_reason.values())):.0f}%" for h, c in by_reason.most_common(6)))
        top = sorted(range(len(rows)), key=lambda k: -int(rows[k][isamp]))[:15]
        for k in sorted(top):
            row = rows[k]
            st = sorted([(int(row[j]), h[6:]) for j, h in stalls if int(row[j]) > 0], reverse=True)[:2]
            lines_out.append(f"   #{k:<5d} {100.0 * int(row[isamp]) / total:5.1f}%  exec {row[iex]:>9s}  {row[ia].strip()[:64]:64s} {st}")
        lines_out.append("")
    open(dst, "w").write("\n".join(lines_out) + "\n")
    print(f"{len(kernels)} kernels -> {dst}")


def _peak():
    """HBM peak the fractions are quoted against: MEASURED_PEAKS.json (driver-written) if present, else the 6.65 TB/s fallback."""
    import json, os
    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


PEAK_GBPS = _peak()


def roofline(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tsc = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}
    out = [["kernel", "launch_us", "grid", "block", "regs", "dram_read_MB", "dram_write_MB", "dram_GBps", f"frac_of_{PEAK_GBPS:.1f}GBps",
            "warps_active_pct", "sm_throughput_pct", "dram_throughput_pct"]]
    for d in data:
        t = float(d[ix["gpu__time_duration.sum"]]) * tsc[units[ix["gpu__time_duration.sum"]]]
        r = float(d[ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]]
        w = float(d[ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]]
        name = d[ix["Kernel Name"]].replace("void ", "").split("(")[0]
        out.append([name, f"{t * 1e6:.1f}", d[ix["launch__grid_size"]], d[ix["launch__block_size"]], d[ix["launch__registers_per_thread"]],
                    f"{r / 1e6:.2f}", f"{w / 1e6:.2f}", f"{(r + w) / t / 1e9:.0f}", f"{(r + w) / t / 1e9 / PEAK_GBPS:.4f}",
                    f"{float(d[ix['sm__warps_active.avg.pct_of_peak_sustained_active']]):.1f}",
                    f"{float(d[ix['sm__throughput.avg.pct_of_peak_sustained_elapsed']]):.1f}",
                    f"{float(d[ix['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]):.1f}"])
    csv.writer(open(dst, "w", newline="")).writerows(out)
    print(f"{len(data)} launches -> {dst}")


if __name__ == "__main__":
    if len(sys.argv) not in (4, 5) or sys.argv[1] not in ("launches", "full", "roofline", "hotspots"):
        sys.exit(__doc__)
    if sys.argv[1] == "hotspots":
        hotspots(sys.argv[2], sys.argv[3], sys.argv[4].split(","))
    elif sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3], tuple(sys.argv[4].split(",")) if len(sys.argv) == 5 else ())
    else:
        {"launches": launches, "roofline": roofline}[sys.argv[1]](sys.argv[2], sys.argv[3])
