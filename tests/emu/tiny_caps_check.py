"""Run in a fresh process with THB_TINY_CAPS=1 (tests/test_emu.py): every growable device structure -- junction / deletion hash
sets, insertion and fusion record buffers, window / indel / fusion task queues, chain queues, joined-record buffer -- starts at
64 entries, so the overflow -> grow -> repeat-the-scan paths run many times; the results must still equal the oracle's and the
join must still equal the committed reference records."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]
assert os.environ.get("THB_TINY_CAPS") == "1"
import build_emu  # noqa: E402
from tophat_b200 import capi, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402
import helpers  # noqa: E402

capi._lib = capi.load_library(build_emu.build())
wl = synth.generate(synth.SynthConfig(contig_lens=(200_000, 80_000), n_pairs=1500, seed=351, indel_prob=0.4, fusion_frac=0.15))
P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20, fusion_search=1, fusion_min_dist=20000)
batches = helpers.pack_both(wl, P)
got, t = helpers.gpu_segjuncs(P, wl.ref, batches)
want, cnt = pyoracle.segjuncs(P, wl.ref, batches)
helpers.assert_same_results(got, want, "tiny capacities")
assert (t.n_windows, t.n_indel_tasks, t.n_rescue_tasks, t.n_juncs_emitted, t.n_fusion_tasks) == \
    (cnt.n_windows, cnt.n_indel_tasks, cnt.n_rescue_tasks, cnt.n_juncs_emitted, cnt.n_fusion_tasks), "task counters are rolled back on a repeated scan"
assert t.kernel_launches > 4, "the scans were expected to be repeated after growth (%d launches)" % t.kernel_launches
assert len(got.junctions) > 64 and len(got.deletions) > 64 and len(got.insertions) > 64 and len(got.fusions) > 64
for name in helpers.join_golden_cases():
    helpers.check_join_golden(name)
import numpy as np  # noqa: E402
import test_flank  # noqa: E402
fc = test_flank.flank_case(11, 2, np.asarray([0, 25, 50, 75, 101]), 120)
nf, _ = test_flank.check_flank(fc, 2, 40)       # the placement append buffer starts at 64 records
assert nf > 64
print("tiny caps ok: %d launches for 2 batches; %d junctions, %d deletions, %d insertions, %d fusions" % (
    t.kernel_launches, len(got.junctions), len(got.deletions), len(got.insertions), len(got.fusions)))
