"""Joins an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv --kernel-name ... ) with nvdisasm -g line info of
the same cubin and prints stall samples aggregated per CUDA source line.
usage: sass_lines.py <ncu_source.csv> <all.sass from nvdisasm -g -c> <mangled-name substring> [top N]"""
import csv, re, sys, collections

ncu_csv, sass, sub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
# 1. instruction -> (file, line) from nvdisasm
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and sub in l and l.rstrip().endswith(":"))
instr_line = []
cur = None
for l in lines[start + 1:]:
    if l.startswith("//--------------------- .text.") or l.startswith("\t.section"):
        if instr_line:
            break
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        instr_line.append(cur)
rows = list(csv.reader(open(ncu_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; body = []
for r in rows[h + 1:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break                                  # next kernel's table
    body.append(r)
si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed")
lsb = hdr.index("stall_long_sb")
print("ncu instructions %d, nvdisasm instructions %d" % (len(body), len(instr_line)))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = 0
for k, r in enumerate(body):
    key = instr_line[k] if k < len(instr_line) else None
    a = agg[key]; a[0] += int(r[si]); a[1] += int(r[ie]); a[2] += int(r[te]); a[3] += int(r[lsb]); tot += int(r[si])
src_cache = {}
def src(key):
    if not key: return ""
    f, n = key
    import glob
    if f not in src_cache:
        c = glob.glob("/root/repo/tophat_b200/csrc/" + f)
        src_cache[f] = open(c[0]).read().split("\n") if c else []
    s = src_cache[f]
    return s[n - 1].strip()[:110] if 0 < n <= len(s) else ""
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%%  samples %6d  long_sb %6d  inst %9d  thr/inst %4.1f  %s:%s  %s" % (
        100.0 * a[0] / max(tot, 1), a[0], a[3], a[1], a[2] / max(a[1], 1), key[0] if key else "?", key[1] if key else "?", src(key)))
