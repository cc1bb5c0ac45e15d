"""Both stages once on device-resident data (for ncu captures of the join kernels): python scripts/prof_join.py [pairs]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from tophat_b200 import capi, synth
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
wl = bench.make_workload(pairs, 0, os.cpu_count() or 1, keep_candidates=True); batches = bench.pack(wl)
P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
ctx = capi.Context(0); ctx.ref_upload(wl.ref)
ctx.segjuncs_begin(P)
for b in batches: ctx.segjuncs_submit(b)
res = ctx.segjuncs_finish()
js = capi.join_sets_from_results(res)
jb = [synth.pack_join_side(wl, wl.left, res.junctions), synth.pack_join_side(wl, wl.right, res.junctions)]
keep, structs = [], []
for b in jb:
    t = {k: torch.from_numpy(np.ascontiguousarray(getattr(b, k)).view(np.uint8).reshape(-1)).cuda() for k in ("bundles", "seg_count", "reads", "hits", "ops_ext")}
    keep.append(t); bc = capi.join_batch_c(b)
    for k in t: setattr(bc, k, t[k].data_ptr())
    structs.append(bc)
torch.cuda.synchronize()
for _ in range(3):
    ctx.join_begin(P, js[0], js[1])
    n = sum(ctx.join_submit_device(bc) for bc in structs)
    t = ctx.join_timing()
    print("join: enum %.3f ms merge %.3f ms (simple %.3f abutting %.3f general %.3f), %d chains %d closures %d joined" % (
        t.enum_ms, t.merge_ms, t.merge_simple_ms, t.merge_abutting_ms, t.merge_general_ms, t.n_chains, t.n_closures, n), flush=True)
