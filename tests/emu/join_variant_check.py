"""Prints a digest of the join's records for a multi-hit-heavy workload (decoy rate 3: reads with more than 8 hits in a segment and
more than 4 chains exercise the tile kernel's generic walk and its parking overflow).  tests/test_emu.py runs it twice -- tile kernel
and, with THB_JOIN_LEGACY=1, the queue kernels -- and compares; both are separately checked against the reference binary elsewhere."""
import hashlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]
import numpy as np  # noqa: E402
if os.environ.get("THB_CHECK_GPU") != "1":
    import build_emu  # noqa: E402
from tophat_b200 import capi, synth  # noqa: E402
import helpers  # noqa: E402

if os.environ.get("THB_CHECK_GPU") != "1":
    capi._lib = capi.load_library(build_emu.build())
h = hashlib.sha256()
tot = 0
for kw in (dict(contig_lens=(300_000, 120_000), n_pairs=2500, seed=881, decoy_rate=3.0, indel_prob=0.2, keep_candidates=True),
           dict(contig_lens=(200_000,), n_pairs=1500, seed=882, decoy_rate=9.0, keep_candidates=True),
           dict(contig_lens=(200_000,), n_pairs=1500, seed=883, read_len=75, decoy_rate=1.0, keep_candidates=True)):
    wl = synth.generate(synth.SynthConfig(**kw))
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20, max_seg_multihits=60)
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    res, _ = helpers.gpu_segjuncs(P, wl.ref, helpers.pack_both(wl), ctx)
    juncs, ins = capi.join_sets_from_results(res)
    ctx.join_begin(P, juncs, ins)
    for side in (wl.left, wl.right):
        jb = synth.pack_join_side(wl, side, res.junctions)
        joined = ctx.join_submit(jb)
        rows = sorted((int(r["bundle"]), int(r["ref_id"]), int(r["left"]), int(r["n_ops"]), int(r["flags"]), int(r["mismatches"]), int(r["edit_dist"]),
                       int(r["splice_mms"]), tuple(int(x) for x in r["ops"][:r["n_ops"]])) for r in joined)
        h.update(repr(rows).encode()); tot += len(rows)
    t = ctx.join_timing()
    h.update(repr((int(t.n_chains), int(t.n_joined))).encode())
    ctx.close()
print("JOIN_DIGEST", h.hexdigest(), tot)
