"""One segment_juncs step on device-resident data (for ncu captures): python scripts/prof_step.py [pairs] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from tophat_b200 import capi
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.make_workload(pairs, 0, os.cpu_count() or 1); batches = bench.pack(wl)
P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
ctx = capi.Context(0); ctx.ref_upload(wl.ref)
keep, structs = [], []
for b in batches:
    t = {k: torch.from_numpy(np.ascontiguousarray(getattr(b, k)).view(np.uint8).reshape(-1)).cuda() for k in ("bundles", "seg_count", "reads", "hits", "partner_hits")}
    keep.append(t); bc = capi.batch_c(b)
    bc.bundles, bc.seg_count, bc.reads, bc.hits, bc.partner_hits = (t[k].data_ptr() for k in ("bundles", "seg_count", "reads", "hits", "partner_hits"))
    structs.append(bc)
torch.cuda.synchronize()
for _ in range(steps):
    ctx.segjuncs_begin(P)
    for bc in structs: ctx.segjuncs_submit_device(bc)
    r = ctx.segjuncs_finish(); tm = ctx.timing()
    print("ms:", {k: round(getattr(tm, k + "_ms"), 4) for k in bench.KERNELS}, "juncs", len(r.junctions), flush=True)
