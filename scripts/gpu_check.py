"""First GPU contact: parity of the CUDA path against the CPU oracle on a few workloads + timings."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tophat_b200 import capi, synth
from oracle import pyoracle

def run(cfg, name):
    t = time.time(); wl = synth.generate(cfg); tg = time.time() - t
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    bl = synth.pack_side(wl.left, wl.right, False)
    br = synth.pack_side(wl.right, wl.left, True, order_base=bl.n_bundles)
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    res = {}
    for rep in range(3):
        ctx.segjuncs_begin(P)
        t = time.time(); ctx.segjuncs_submit(bl); ctx.segjuncs_submit(br); t1 = time.time() - t
        got = ctx.segjuncs_finish(); tm = ctx.timing()
        res = dict(name=name, gen_s=tg, bundles=bl.n_bundles + br.n_bundles, submit_s=t1, scan_ms=tm.scan_kernel_ms,
                   h2d_ms=tm.h2d_ms, finish_ms=tm.finish_ms, launches=tm.kernel_launches, windows=tm.n_windows,
                   indel=tm.n_indel_tasks, rescue=tm.n_rescue_tasks, emits=tm.n_juncs_emitted, alg_bytes=tm.algorithmic_bytes,
                   juncs=len(got.junctions), dels=len(got.deletions), ins=len(got.insertions))
    t = time.time(); want, cnt = pyoracle.segjuncs(P, wl.ref, [bl, br]); res['oracle_s'] = time.time() - t
    ok = True
    for k in ("junctions", "deletions", "insertions"):
        a, b = getattr(got, k), getattr(want, k)
        same = a.shape == b.shape and bool((a == b).all())
        res['eq_' + k] = same; ok &= same
    res['counters_eq'] = (tm.n_windows == cnt.n_windows, tm.n_indel_tasks == cnt.n_indel_tasks, tm.n_rescue_tasks == cnt.n_rescue_tasks, tm.n_juncs_emitted == cnt.n_juncs_emitted)
    res['gbps'] = res['alg_bytes'] / (res['scan_ms'] * 1e-3) / 1e9 if res['scan_ms'] else 0
    print(json.dumps(res), flush=True)
    ctx.close()
    return ok

if __name__ == "__main__":
    ok = True
    ok &= run(synth.SynthConfig(contig_lens=(600_000, 200_000), n_pairs=4000, indel_prob=0.25, seed=7), "tiny")
    ok &= run(synth.SynthConfig(contig_lens=(8_000_000, 3_000_000), n_pairs=200_000, seed=11), "200k")
    ok &= run(synth.SynthConfig(contig_lens=(5_000_000,), n_pairs=100_000, indel_prob=0.5, seed=13), "indel100k")
    if len(sys.argv) > 1:
        ok &= run(synth.SynthConfig(contig_lens=(64_444_167,), n_pairs=int(sys.argv[1]), seed=17), "big")
    print("ALL OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
