"""Run in a fresh process with THB_TINY_CAPS=1 ON THE GPU (tests/test_gpu_scale.py): every growable device structure --
junction / deletion hash sets, insertion and fusion record buffers, window / indel / fusion task queues, the chain queue, the
joined-record buffer -- starts at 64 entries, so the overflow -> grow -> repeat-the-scan paths of thb_api.cu run many times on
the real kernels.  The results must still equal the oracle's and the join must still equal the committed reference records."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, HERE]
assert os.environ.get("THB_TINY_CAPS") == "1"
from tophat_b200 import capi, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402
import helpers  # noqa: E402

for kw, over in [
    (dict(contig_lens=(200_000, 80_000), n_pairs=1500, seed=351, indel_prob=0.4, fusion_frac=0.15), dict(fusion_search=1, fusion_min_dist=20000)),
    (dict(contig_lens=(2_000_000, 700_000), n_pairs=60_000, seed=352, indel_prob=0.3), {}),
]:
    wl = synth.generate(synth.SynthConfig(**kw), workers=4)
    o = dict(inner_dist_mean=50, inner_dist_std_dev=20); o.update(over)
    P = capi.default_params(**o)
    batches = helpers.pack_both(wl, P)
    got, t = helpers.gpu_segjuncs(P, wl.ref, batches)
    want, cnt = pyoracle.segjuncs(P, wl.ref, batches)
    helpers.assert_same_results(got, want, "tiny capacities %r" % (kw,))
    assert (t.n_windows, t.n_indel_tasks, t.n_rescue_tasks, t.n_juncs_emitted, t.n_fusion_tasks) == \
        (cnt.n_windows, cnt.n_indel_tasks, cnt.n_rescue_tasks, cnt.n_juncs_emitted, cnt.n_fusion_tasks), "task counters are rolled back on a repeated scan"
    assert t.kernel_launches > 4, "the scans were expected to be repeated after growth (%d launches)" % t.kernel_launches
    print("tiny caps ok: %d scan launches for 2 batches; %d junctions, %d deletions, %d insertions, %d fusions" % (
        t.kernel_launches, len(got.junctions), len(got.deletions), len(got.insertions), len(got.fusions)))
for name in helpers.join_golden_cases():
    helpers.check_join_golden(name)
print("tiny caps ok: join goldens")
import numpy as np  # noqa: E402
import test_flank  # noqa: E402
for nm in ("v2_101bp", "v3_two_word_contigs"):
    seed, v, bounds, n_reads, max_hits, npol, kw = test_flank.FLANK_CASES[nm]
    n, _ = test_flank.check_flank(test_flank.flank_case(seed, v, np.asarray(bounds), n_reads, **kw), v, max_hits, npol)
    assert n > 40
print("tiny caps ok: flank matcher (append buffer growth, grid-stride rounds)")
