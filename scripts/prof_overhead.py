"""Wall-clock split of one bench step (device-resident) to find host-side overheads: python scripts/prof_overhead.py [pairs]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from tophat_b200 import capi, synth
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
wl = bench.make_workload(pairs, 0, os.cpu_count() or 1, keep_candidates=True); batches = bench.pack(wl)
P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
ctx = capi.Context(0); ctx.ref_upload(wl.ref)
def dev(bl, fields, mk):
    keep, st = [], []
    for b in bl:
        t = {k: torch.from_numpy(np.ascontiguousarray(getattr(b, k)).view(np.uint8).reshape(-1)).cuda() for k in fields}
        keep.append(t); bc = mk(b)
        for k in t: setattr(bc, k, t[k].data_ptr())
        st.append(bc)
    return keep, st
k1, s1 = dev(batches, ("bundles", "seg_count", "reads", "hits", "partner_hits"), capi.batch_c)
ctx.segjuncs_begin(P); [ctx.segjuncs_submit_device(x) for x in s1]; res = ctx.segjuncs_finish()
js = capi.join_sets_from_results(res)
jb = [synth.pack_join_side(wl, wl.left, res.junctions), synth.pack_join_side(wl, wl.right, res.junctions)]
k2, s2 = dev(jb, ("bundles", "seg_count", "reads", "hits", "ops_ext"), capi.join_batch_c)
torch.cuda.synchronize()
for rep in range(4):
    T = {}
    def tick(name, f):
        t = time.perf_counter(); r = f(); torch.cuda.synchronize(); T[name] = T.get(name, 0) + (time.perf_counter() - t) * 1e3; return r
    tick("begin", lambda: ctx.segjuncs_begin(P))
    for x in s1: tick("submit", lambda: ctx.segjuncs_submit_device(x))
    tick("finish", lambda: ctx.segjuncs_finish())
    tm = ctx.timing()
    tick("join_begin", lambda: ctx.join_begin(P, js[0], js[1]))
    for x in s2: tick("join_submit", lambda: ctx.join_submit_device(x))
    jt = ctx.join_timing()
    print({k: round(v, 3) for k, v in T.items()}, "total %.3f" % sum(T.values()), "| kernels: scan %.3f finish %.3f join %.3f" % (tm.scan_kernel_ms, tm.finish_ms, jt.kernel_ms), flush=True)
