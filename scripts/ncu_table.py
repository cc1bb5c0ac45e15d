"""Per-launch table of the metrics the profiles/ summaries quote, from an .ncu-rep (ncu -i ... --page raw --csv).
usage: python scripts/ncu_table.py <report.ncu-rep> [> table.md]"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time us", 1.0), ("smsp__inst_executed.sum", "warp inst (M)", 1e-6),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("dram__bytes_read.sum", "DRAM rd", None), ("dram__bytes_write.sum", "DRAM wr", None),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0), ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1.0),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb", 1.0),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_inst", 1.0),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait", 1.0),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch", 1.0)]
UNIT = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    print("| # | kernel | grid x block | " + " | ".join(c[1] for c in COLS) + " | DRAM GB/s |")
    print("|---|---|---|" + "---|" * (len(COLS) + 1))
    for n, d in enumerate(data):
        name = d[ik].split("(")[0].replace("void ", "").replace("thb::", "")
        vals = []; tus = None; dram = 0.0
        for key, label, scale in COLS:
            if key not in hdr:
                vals.append("-"); continue
            i = hdr.index(key); v = float(d[i].replace(",", ""))
            if scale is None:
                v *= UNIT.get(units[i], 1.0); dram += v; vals.append("%.3f GB" % v)
            else:
                if key == "gpu__time_duration.sum":
                    v *= {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[i], 1.0); tus = v
                vals.append(("%.1f" if abs(v * scale) >= 10 else "%.2f") % (v * scale))
        g = d[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "?"; b = d[hdr.index("launch__block_size")] if "launch__block_size" in hdr else "?"
        print("| %d | %s | %s x %s | %s | %.0f |" % (n, name, g, b, " | ".join(vals), dram / (tus * 1e-6) if tus else 0.0))


if __name__ == "__main__":
    main()
