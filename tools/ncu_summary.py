#!/usr/bin/env python
"""Condense `ncu -i rep --page raw --csv` (one row per profiled launch, ~2000 columns) into the columns the profiles/
summaries carry, plus achieved DRAM GB/s and its fraction of the measured HBM peak.

  python tools/ncu_summary.py raw.csv [out.csv]
"""
import csv
import json
import os
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_unit(value, unit, want):
    scale = {"": 1.0, "K": 1e3, "M": 1e6, "G": 1e9, "T": 1e12, "n": 1e-9, "u": 1e-6, "m": 1e-3}
    v = float(value.replace(",", ""))
    if want == "s":   # durations: ns / us / ms / s
        pre = unit[:-1] if unit.endswith("s") and unit not in ("s",) else ""
        return v * scale.get(pre, 1.0)
    if want == "byte":
        pre = unit.replace("byte", "")
        return v * scale.get(pre, 1.0)
    return v


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    out = [["Kernel Name", "ms", "dram_read_GB", "dram_write_GB", "dram_GBps", "frac_of_hbm_peak"] + COLS[3:]]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        t = to_unit(r[idx[COLS[0]]], units[idx[COLS[0]]], "s")
        rd = to_unit(r[idx[COLS[1]]], units[idx[COLS[1]]], "byte")
        wr = to_unit(r[idx[COLS[2]]], units[idx[COLS[2]]], "byte")
        gbps = (rd + wr) / t / 1e9
        line = [r[idx["Kernel Name"]], f"{t * 1e3:.4f}", f"{rd / 1e9:.4f}", f"{wr / 1e9:.4f}", f"{gbps:.1f}", f"{gbps / peak:.3f}"]
        for c in COLS[3:]:
            line.append(r[idx[c]] if c in idx else "")
        out.append(line)
    w = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
    w.writerows(out)


if __name__ == "__main__":
    main()
