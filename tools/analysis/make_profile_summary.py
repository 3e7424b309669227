#!/usr/bin/env python
"""Build profiles/<name>.md from `ncu -i X.ncu-rep --page raw --csv` exports: per profiled launch the counters BASELINE.json's
north_star asks for (DRAM bytes, achieved HBM GB/s against the measured peak, SM / LSU / pipe utilisation, issue activity).
usage: make_profile_summary.py title out.md raw1.csv [raw2.csv ...]"""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
COLS = [("gpu__time_duration.sum", "time us", 1.0), ("dram__bytes_read.sum", "DRAM rd MB", 1.0), ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %", 1.0), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %", 1.0),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %", 1.0), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 %", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1.0), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1.0),
        ("smsp__inst_executed.sum", "warp-instr M", 1e-6), ("launch__grid_size", "grid", 1.0), ("launch__registers_per_thread", "regs", 1.0)]


def main():
    title, out = sys.argv[1], sys.argv[2]
    peak = 6547.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    lines = [f"# {title}", "", f"`ncu --set full --clock-control none` (cold caches, serialised launches: compare shares, not absolutes); HBM peak = {peak} GB/s "
             "(measured copy bandwidth, MEASURED_PEAKS.json).  `HBM GB/s` = (DRAM read + write) / time; `of peak` = that / the measured peak.", ""]
    hdr_line = "| kernel | " + " | ".join(c[1] for c in COLS[:3]) + " | HBM GB/s | of peak | " + " | ".join(c[1] for c in COLS[3:]) + " |"
    lines += [hdr_line, "|---|" + "---:|" * (len(COLS) + 2)]
    for path in sys.argv[3:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        for r in rows[2:]:
            name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
            v = []
            for key, _, scale in COLS:
                if key in hdr:
                    x = float(r[hdr.index(key)].replace(",", "")) * scale
                    u = units[hdr.index(key)]
                    if key.startswith("dram__bytes") and u == "Kbyte":
                        x /= 1e3
                    if key.startswith("dram__bytes") and u == "byte":
                        x /= 1e6
                    if key.startswith("dram__bytes") and u == "Gbyte":
                        x *= 1e3
                    if key == "gpu__time_duration.sum" and u == "ms":
                        x *= 1e3
                    if key == "gpu__time_duration.sum" and u in ("ns", "nsecond"):
                        x /= 1e3
                    v.append(x)
                else:
                    v.append(float("nan"))
            gbs = (v[1] + v[2]) * 1e6 / (v[0] * 1e-6) / 1e9 if v[0] else 0.0
            cells = [f"{v[0]:.1f}", f"{v[1]:.2f}", f"{v[2]:.2f}", f"{gbs:.0f}", f"{gbs / peak:.3f}"] + [f"{x:.1f}" if i < 9 else (f"{x:.1f}" if COLS[3 + i][0].startswith("smsp__inst") else f"{x:.0f}") for i, x in enumerate(v[3:])]
            lines.append(f"| {name} | " + " | ".join(cells) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
