#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) as a markdown table: one row per kernel launch.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/rNN_x_summary.md]
"""
import csv
import io
import re
import subprocess
import sys

COLS = [
    ("us", "gpu__time_duration.sum", 1.0),
    ("grid", "launch__grid_size", 1.0),
    ("blk", "launch__block_size", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("smemKB", "launch__shared_mem_per_block_dynamic", 1.0),
    ("waves", "launch__waves_per_multiprocessor", 1.0),
    ("dramRdMB", "dram__bytes_read.sum", 1.0),
    ("dramWrMB", "dram__bytes_write.sum", 1.0),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("l2MB", "lts__t_bytes.sum", 1.0),
    ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("issue%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("instM", "smsp__inst_executed.sum", 1e-6),
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio")


def to_float(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def unit_scale(unit, want):
    table = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6} if want == "us" else \
        {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    return table.get(unit, 1.0)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx["Kernel Name"]
    stalls = [(m.group(1), i) for i, h in enumerate(hdr) if (m := STALL.match(h))]
    print("| # | kernel | " + " | ".join(c[0] for c in COLS) + " | top stalls (warps per issue) |")
    print("|" + "---|" * (len(COLS) + 3))
    for n, r in enumerate(data):
        cells = []
        for label, metric, sc in COLS:
            if metric not in idx:
                cells.append("-")
                continue
            v = to_float(r[idx[metric]]) * sc
            u = units[idx[metric]]
            if label == "us":
                v *= unit_scale(u, "us")
            elif label.endswith("MB") and label != "smemKB":
                v *= unit_scale(u, "MB")
            cells.append(f"{v:.1f}" if abs(v) < 1e5 else f"{v:.3g}")
        st = sorted(((to_float(r[i]), nm) for nm, i in stalls), reverse=True)[:4]
        name = re.sub(r"\(.*", "", r[name_i])
        name = re.sub(r"void |tcg::|tc_gemm_kernel", "", name)[:70]
        print(f"| {n} | {name} | " + " | ".join(cells) + " | " + ", ".join(f"{nm} {v:.1f}" for v, nm in st) + " |")


if __name__ == "__main__":
    main()
