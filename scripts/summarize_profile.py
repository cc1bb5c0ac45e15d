"""Summarises ncu outputs brought back from gpurun into small tracked files under profiles/.

  python scripts/summarize_profile.py <tag>      reads gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep,
                                                 gpurun_out/bench_<tag>.json; writes profiles/<tag>_*.{csv,md,json}
"""
import csv, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles"); os.makedirs(P, exist_ok=True)
out = ["# ncu summary `%s`" % tag, ""]

lp = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if not l.startswith("=="))]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv: continue
        v = float(r[iv].replace(",", "")); u = r[iu]
        v_us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0 if u in ("ms", "msecond") else v)
        name = r[ik].split("(")[0]
        d = per.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += v_us
    tot = sum(d[1] for d in per.values())
    out += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, bench.py --steps 2 --warmup 1; cold-cache, serialised)", "",
            "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f%% |" % (k, n, t, 100 * t / tot if tot else 0))
    out.append("")
    with open(os.path.join(P, "%s_launches.csv" % tag), "w") as f:
        f.write("".join(l for l in open(lp) if not l.startswith("==")))

rp = os.path.join(G, "prof_%s.ncu-rep" % tag)
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
    out += ["## `ncu --set full` captures", ""]
    traffic = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        name = name[5:] if name.startswith("void ") else name
        name = name.split("<")[0].replace("thb::", "")
        out += ["### `%s`" % name, "", "| metric | value | unit |", "|---|---|---|"]
        vals = {}
        for w in want:
            if w in hdr:
                i = hdr.index(w); out.append("| %s | %s | %s |" % (w, r[i], units[i])); vals[w] = (r[i], units[i])
        out.append("")
        def tobytes(v, u):
            v = float(v.replace(",", "")); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        if "dram__bytes_read.sum" in vals:
            traffic.setdefault(name, []).append(tobytes(*vals["dram__bytes_read.sum"]) + tobytes(*vals["dram__bytes_write.sum"]))
    tj = {k + "_dram_bytes_per_launch": sum(v) / len(v) for k, v in traffic.items()}
    tj["source"] = "profiles/%s_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % tag
    json.dump(tj, open(os.path.join(P, "traffic.json"), "w"), indent=1)

bp = os.path.join(G, "bench_%s.json" % tag)
if os.path.exists(bp):
    out += ["## bench.py line of the same session", "", "```", open(bp).read().strip(), "```", ""]
open(os.path.join(P, "%s_summary.md" % tag), "w").write("\n".join(out))
print("\n".join(out[:60]))
