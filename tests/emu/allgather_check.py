"""Run in a fresh process by tests/test_emu.py: thb_segjuncs_allgather at world size 1 (fake libnccl, emulated kernels)
must leave every result set unchanged -- junction / deletion keys, insertion records and fusion records all travel
through the send / receive staging areas and are re-inserted."""
import ctypes
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]
import build_emu  # noqa: E402

lib = build_emu.build()
fake = os.path.join(build_emu.OUT_DIR, "libnccl.so.2")
if not os.path.exists(fake) or os.path.getmtime(fake) < os.path.getmtime(os.path.join(HERE, "fake_nccl.c")):
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-Wl,-soname,libnccl.so.2", "-o", fake, os.path.join(HERE, "fake_nccl.c")], check=True)
ctypes.CDLL(fake, mode=ctypes.RTLD_GLOBAL)          # dlopen("libnccl.so.2") inside the library now resolves to it

from tophat_b200 import capi, synth  # noqa: E402
import helpers  # noqa: E402

capi._lib = capi.load_library(lib)
wl = synth.generate(synth.SynthConfig(contig_lens=(200_000, 80_000), n_pairs=2000, seed=341, indel_prob=0.4, fusion_frac=0.2))
P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20, fusion_search=1, fusion_min_dist=20000)
batches = helpers.pack_both(wl, P)
plain, _ = helpers.gpu_segjuncs(P, wl.ref, batches)
ctx = capi.Context(0); ctx.ref_upload(wl.ref)
ctx.comm_init(ctx.nccl_unique_id(), 0, 1)
ctx.segjuncs_begin(P)
for b in batches:
    ctx.segjuncs_submit(b)
ctx.segjuncs_allgather()
got = ctx.segjuncs_finish()
helpers.assert_same_results(got, plain, "all-gather at world size 1")
assert len(got.fusions) > 50 and len(got.insertions) > 20 and len(got.deletions) > 20
ctx.close()
# "world size 3" of the stand-in: every peer sends what this rank sends -- the union must still be the single-rank sets (keys
# de-duplicated by the sort + unique of thb_segjuncs_finish, insertions by first-wins), fusion counts add up
ctx = capi.Context(0); ctx.ref_upload(wl.ref)
ctx.comm_init(ctx.nccl_unique_id(), 0, 3)
ctx.segjuncs_begin(P)
for b in batches:
    ctx.segjuncs_submit(b)
ctx.segjuncs_allgather()
try:
    ctx.segjuncs_submit(batches[0]); raise SystemExit("a submit after the all-gather must be refused")
except capi.ThbError:
    pass
got3 = ctx.segjuncs_finish()
for name in ("junctions", "deletions", "insertions"):
    a, b = getattr(got3, name), getattr(plain, name)
    assert a.shape == b.shape and (a == b).all(), "replicated all-gather: %s differ" % name
assert len(got3.fusions) == len(plain.fusions)
for f in ("ref_id1", "ref_id2", "left", "right", "dir", "edit_dist"):
    assert (got3.fusions[f] == plain.fusions[f]).all(), f
assert (got3.fusions["count"] == 3 * plain.fusions["count"]).all()
# the resident hand-off after an all-gather uses the gathered union
P2 = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
ctx.segjuncs_begin(P2)
for b in helpers.pack_both(wl, P2):
    ctx.segjuncs_submit(b)
ctx.segjuncs_allgather()
raw = ctx.segjuncs_finish_resident(); ctx.join_begin_resident(P2); ctx.segjuncs_fetch()
assert raw.n_junctions == len(helpers.gpu_segjuncs(P2, wl.ref, helpers.pack_both(wl, P2))[0].junctions)
ctx.close()
print("allgather ok: %d junctions, %d deletions, %d insertions, %d fusions" % (len(got.junctions), len(got.deletions), len(got.insertions), len(got.fusions)))
