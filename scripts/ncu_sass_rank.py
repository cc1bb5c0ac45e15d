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
iS = hdr.index("Source"); iSec = hdr.index("L2 Theoretical Sectors Global"); iId = hdr.index("L2 Theoretical Sectors Global Ideal"); iLoc = hdr.index("L2 Theoretical Sectors Local")
iAcc = hdr.index("Access Operation"); iSz = hdr.index("Access Size"); iIE = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); iSm = hdr.index("# Samples")
for k in secs[:1]:
    data = k["data"]
    print(k["name"][:100], len(data), "instructions")
    tot = sum(int(r[iSec]) for r in data); totl = sum(int(r[iLoc]) for r in data); ti = sum(int(r[iIE]) for r in data); ts = sum(int(r[iSm]) for r in data)
    print("total global sectors", tot, "= %.1f MB" % (tot * 32 / 1e6), "local sectors", totl, "= %.1f MB" % (totl * 32 / 1e6), "inst", ti, "samples", ts)
    mem = [(int(r[iSec]) + int(r[iLoc]), int(r[iId]), i, r) for i, r in enumerate(data) if int(r[iSec]) > 0 or int(r[iLoc]) > 0]
    mem.sort(key=lambda x: -x[0])
    for sec, ideal, i, r in mem[:top]:
        print(i, r[iS].strip()[:50].ljust(50), r[iAcc][:5], r[iSz], "sec", sec, "ideal", ideal, "inst", r[iIE], "thr", r[iT], "loc", r[iLoc], "smp", r[iSm])
