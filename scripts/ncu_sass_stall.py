import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "data": []}; secs.append(cur)
    elif r and r[0] == "Address":
        hdr = r
    elif r and r[0].startswith("0x") and len(r) == len(hdr):
        cur["data"].append(r)
iS = hdr.index("Source"); iSm = hdr.index("# Samples"); iIE = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = secs[0]["data"]
ts = sum(int(r[iSm]) for r in data)
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stalls}
print("samples", ts, {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > ts * 0.01})
rank = sorted(range(len(data)), key=lambda i: -int(data[i][iSm]))
for i in rank[:top]:
    r = data[i]
    st = {hdr[j][6:]: int(r[j]) for j in stalls if int(r[j]) > 0.15 * int(r[iSm])}
    print(i, r[iS].strip()[:50].ljust(50), "smp", r[iSm], "inst", r[iIE], "thr/inst %.1f" % (int(r[iT]) / max(1, int(r[iIE]))), st)
