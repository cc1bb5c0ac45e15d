"""GPU parity at the shapes the metric is quoted on (BASELINE configs[2]: hg38-sized reference) and of the paths that only
run when a device structure overflows.  Through the C ABI, against the CPU oracle and the reference's own binary."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import helpers
from tophat_b200 import capi, shard, synth
from oracle import pyoracle

pytestmark = pytest.mark.gpu
SMALL = os.environ.get("THB_TEST_EMU") == "1"       # development aid (pytest --emu): the emulated kernels get a tenth of the reads

# the 24 primary contigs of hg38 (chr1..22, X, Y), 3.09 Gbp
HG38 = (248_956_422, 242_193_529, 198_295_559, 190_214_555, 181_538_259, 170_805_979, 159_345_973, 145_138_636, 138_394_717,
        133_797_422, 135_086_622, 133_275_309, 114_364_328, 107_043_718, 101_991_189, 90_338_345, 83_257_441, 80_373_285,
        58_617_616, 64_444_167, 46_709_983, 50_818_468, 156_040_895, 57_227_415)


def test_hg38_sized_reference_matches_oracle():
    """The full-size image: 24 contigs, 3.09 Gbp, 0.5 % of the bases in N runs, global coordinates up to 3.1e9 (above 2^31), 1.2 GB of
    bit planes that no longer sit in L2.  120 k pairs -> every set and every task counter equal to the CPU oracle's; then the same
    input in three shards submitted out of order."""
    wl = synth.generate(synth.SynthConfig(contig_lens=HG38, n_pairs=12_000 if SMALL else 120_000, seed=20240611, indel_prob=0.2, chunk=30_000, keep_candidates=True), workers=min(4, os.cpu_count() or 1))
    assert int(wl.ref.contig_start[-1]) > (1 << 31)
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    batches = helpers.pack_both(wl)
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    got, t = helpers.gpu_segjuncs(P, wl.ref, batches, ctx)
    want, cnt = pyoracle.segjuncs(P, wl.ref, batches)
    helpers.assert_same_results(got, want, "hg38-sized")
    assert (t.n_windows, t.n_indel_tasks, t.n_rescue_tasks, t.n_juncs_emitted) == (cnt.n_windows, cnt.n_indel_tasks, cnt.n_rescue_tasks, cnt.n_juncs_emitted)
    assert len(got.junctions) > (2_000 if SMALL else 20_000) and len(set(got.junctions["ref_id"].tolist())) == 24
    assert (got.junctions["ref_id"] >= 20).sum() > (10 if SMALL else 100)                      # junctions beyond global base 2^31
    parts = [shard.shard_batch(b, r, 3) for r in (2, 0, 1) for b in batches]
    got3, _ = helpers.gpu_segjuncs(P, wl.ref, parts, ctx)
    helpers.assert_same_results(got3, want, "hg38-sized, 3 shards")
    # stage 2 on the same image: every joined alignment re-read independently (numpy walk of CIGAR over genome and read)
    juncs, ins = capi.join_sets_from_results(got)
    ctx.join_begin(P, juncs, ins)
    n_tot = 0
    for side in (wl.left, wl.right):
        jb = synth.pack_join_side(wl, side, got.junctions)
        joined = ctx.join_submit(jb)
        n_tot += len(joined)
        import test_gpu_join
        mm, ln, nn = test_gpu_join._recount_mismatches(wl, jb, joined[:4000])
        assert (ln == jb.bundles["read_len"][joined["bundle"][:4000]]).all()
        assert ((mm == joined["mismatches"][:4000]) | (mm + nn == joined["mismatches"][:4000])).all()
    assert n_tot > (6_000 if SMALL else 60_000)
    ctx.close()


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
def test_24_contig_join_matches_reference_binary():
    """hg38's contig layout at 1/64 scale (24 contigs, 48 Mbp, N runs), 100 k pairs: segment.juncs byte-identical to the reference's
    segment_juncs and the packed-batch join record-identical to the reference's long_spanning_reads (both mates)."""
    import test_gpu_join
    lens = tuple(max(200_000, n // 64) for n in HG38)
    wl = synth.generate(synth.SynthConfig(contig_lens=lens, n_pairs=10_000 if SMALL else 100_000, seed=778, indel_prob=0.1, keep_truth=True, chunk=25_000), workers=min(4, os.cpu_count() or 1))
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    res, _ = helpers.gpu_segjuncs(P, wl.ref, helpers.pack_both(wl), ctx)
    juncs, ins = capi.join_sets_from_results(res)
    ctx.join_begin(P, juncs, ins)
    names = wl.ref.names
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as td:
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg)      # -p1: the reference's exact answer
        assert open(outs["juncs"]).read() == pyoracle.format_juncs(res.junctions, names)
        assert open(outs["deletions"]).read() == pyoracle.format_deletions(res.deletions, names)
        assert open(outs["insertions"]).read() == pyoracle.format_insertions(res.insertions, names)
        jin = pyoracle.make_join_inputs(wl, files, outs, td, nseg, fast=True)
        for sname, side in (("left", wl.left), ("right", wl.right)):
            batch = synth.pack_join_side(wl, side, res.junctions)
            joined = ctx.join_submit(batch)
            got = test_gpu_join.joined_to_keys(joined, batch, P)
            ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg,
                                                       side=sname, tag=".ref")
            _, recs = pyoracle.read_bam(ref_bam)
            want = set((int(r[0]), names.index(r[2]) + 1, r[3], r[5], r[1], r[11]["NM"]) for r in recs)
            assert got == want, "%s: %d vs %d alignments; only ours %r; only reference %r" % (
                sname, len(got), len(want), sorted(got - want)[:3], sorted(want - got)[:3])
            assert len(want) > (2_500 if SMALL else 25_000)
    ctx.close()


@pytest.mark.skipif(SMALL, reason="runs the real library in a fresh process")
def test_growth_paths_on_hardware():
    """THB_TINY_CAPS=1 in a fresh process on the GPU: overflow -> grow -> repeat of every device structure (tests/gpu_tiny_caps_check.py)."""
    env = dict(os.environ, THB_TINY_CAPS="1")
    r = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "tests", "gpu_tiny_caps_check.py")], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "tiny caps ok: join goldens" in r.stdout


@pytest.mark.skipif(SMALL, reason="runs the real library in a fresh process")
def test_join_tile_kernel_equals_queue_kernels_on_hardware():
    """The tile kernel against round 1's queue kernels (THB_JOIN_LEGACY=1) on multi-hit-heavy reads, on the GPU."""
    outs = []
    for legacy in (False, True):
        env = dict(os.environ, THB_CHECK_GPU="1"); env.pop("THB_JOIN_LEGACY", None); env.pop("THB_JOIN_TILE", None)
        if not legacy:
            env["THB_JOIN_TILE"] = "1"
        r = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "tests", "emu", "join_variant_check.py")], capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith("JOIN_DIGEST")][-1])
    assert outs[0] == outs[1] and int(outs[0].split()[-1]) > 3000
