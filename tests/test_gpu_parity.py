"""GPU parity tests proper: the CUDA path, called through the C ABI (libtophat_b200.so), against the CPU
oracle and the committed outputs of the reference binary.  Integer/byte work: the bar is bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers
import kat
from tophat_b200 import capi, shard, synth
from oracle import pyoracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return capi.load_library()


@pytest.mark.parametrize("name", helpers.golden_cases())
def test_gpu_matches_reference_golden(name):
    wl, P, want = helpers.load_golden(name)
    got, _ = helpers.gpu_segjuncs(P, wl.ref, helpers.pack_both(wl, P))
    txt = helpers.as_text(got, wl.ref.names)
    for k in want:
        assert txt[k] == want[k], "%s: segment.%s differs from the reference binary's output" % (name, k)


@pytest.mark.parametrize("kw,over", [
    (dict(contig_lens=(900_000, 300_000, 50_000), n_pairs=20_000, seed=201), {}),
    (dict(contig_lens=(700_000,), n_pairs=20_000, seed=202, indel_prob=0.5), {}),
    (dict(contig_lens=(700_000,), n_pairs=10_000, seed=203, indel_prob=0.3, n_rate=0.01, sub_rate=0.02), {}),
    (dict(contig_lens=(400_000, 400_000), n_pairs=10_000, seed=204, decoy_rate=3.0), {}),
    (dict(contig_lens=(400_000,), n_pairs=10_000, seed=205, ref_n_frac=0.05), dict(inner_dist_mean=200)),
    (dict(contig_lens=(400_000,), n_pairs=10_000, seed=206, indel_prob=0.2), dict(library_type=3)),
    (dict(contig_lens=(400_000,), n_pairs=10_000, seed=207, indel_prob=0.2), dict(inner_dist_mean=10, inner_dist_std_dev=40)),
    (dict(contig_lens=(400_000,), n_pairs=6_000, seed=208, read_len=75, indel_prob=0.2), {}),
    (dict(contig_lens=(400_000,), n_pairs=6_000, seed=209, read_len=150, indel_prob=0.2), {}),
    (dict(contig_lens=(400_000,), n_pairs=6_000, seed=210, read_len=100, segment_length=20, indel_prob=0.2), dict(segment_length=20)),
])
def test_gpu_matches_oracle(kw, over):
    wl = synth.generate(synth.SynthConfig(**kw))
    o = dict(inner_dist_mean=50, inner_dist_std_dev=20); o.update(over)
    P = capi.default_params(**o)
    batches = helpers.pack_both(wl)
    got, t = helpers.gpu_segjuncs(P, wl.ref, batches)
    want, cnt = pyoracle.segjuncs(P, wl.ref, batches)
    helpers.assert_same_results(got, want, str(kw))
    assert (t.n_windows, t.n_indel_tasks, t.n_rescue_tasks, t.n_juncs_emitted) == \
        (cnt.n_windows, cnt.n_indel_tasks, cnt.n_rescue_tasks, cnt.n_juncs_emitted)
    assert t.kernel_launches >= 1


@pytest.mark.parametrize("kw,over", [
    (dict(contig_lens=(900_000, 300_000, 50_000), n_pairs=20_000, seed=231, indel_prob=0.1, fusion_frac=0.1), dict(fusion_min_dist=50000)),
    (dict(contig_lens=(400_000, 400_000), n_pairs=10_000, seed=232, fusion_frac=0.3, decoy_rate=2.0), {}),
    (dict(contig_lens=(400_000,), n_pairs=6_000, seed=233, read_len=150, fusion_frac=0.2, n_rate=0.005), dict(fusion_anchor_length=30, fusion_min_dist=100)),
    (dict(contig_lens=(400_000, 100_000), n_pairs=6_000, seed=234, read_len=75, fusion_frac=0.2, ref_n_frac=0.05), dict(inner_dist_mean=10, inner_dist_std_dev=40)),
])
def test_gpu_fusions_match_oracle(kw, over):
    """--fusion-search: find_fusions / detect_fusion (segment_juncs.cpp:2976-3291, 2629-2805) -- ff / fr / rf / rr, intra- and
    inter-contig, with the mate-flank rescue; counts and minimum edit distances per fusion included."""
    wl = synth.generate(synth.SynthConfig(**kw))
    o = dict(inner_dist_mean=50, inner_dist_std_dev=20, fusion_search=1); o.update(over)
    P = capi.default_params(**o)
    batches = helpers.pack_both(wl, P)
    got, t = helpers.gpu_segjuncs(P, wl.ref, batches)
    want, cnt = pyoracle.segjuncs(P, wl.ref, batches)
    helpers.assert_same_results(got, want, str(kw))
    assert len(set(int(d) for d in want.fusions["dir"])) == 4 and len(want.fusions) > 200
    assert (t.n_windows, t.n_indel_tasks, t.n_rescue_tasks, t.n_juncs_emitted, t.n_fusion_tasks) == \
        (cnt.n_windows, cnt.n_indel_tasks, cnt.n_rescue_tasks, cnt.n_juncs_emitted, cnt.n_fusion_tasks)
    # fusion records are order-independent: shards in any order give the same reduced set
    parts = [shard.shard_batch(b, r, 3) for r in (1, 2, 0) for b in batches]
    got3, _ = helpers.gpu_segjuncs(P, wl.ref, parts)
    helpers.assert_same_results(got3, want, "3 shards " + str(kw))


@pytest.mark.parametrize("case", ["kat_junction", "kat_junction_seg1_unmapped", "kat_deletion", "kat_insertion", "kat_q0_quirk"])
def test_gpu_known_answers(case):
    contigs, reads, exp = getattr(kat, case)()
    ref, batch = helpers.manual_workload(contigs, reads)
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    got, _ = helpers.gpu_segjuncs(P, ref, [batch])
    want, _ = pyoracle.segjuncs(P, ref, [batch])
    helpers.assert_same_results(got, want, case)
    exp = exp if isinstance(exp, list) else [exp]
    text = pyoracle.format_juncs(got.junctions, ref.names) + pyoracle.format_deletions(got.deletions, ref.names) + \
        pyoracle.format_insertions(got.insertions, ref.names)
    for e in exp:
        assert "\t".join(str(x) for x in e) in text.splitlines()


def test_empty_and_repeated_runs():
    wl, P, want = helpers.load_golden("indel_heavy")
    batches = helpers.pack_both(wl)
    b = batches[0]
    empty = synth.PackedBatch(b.n_segs, b.read_words, b.bundles[:0], b.seg_count[:0], b.reads[:0], b.hits[:0], b.partner_hits[:0])
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    got, _ = helpers.gpu_segjuncs(P, wl.ref, [empty], ctx)
    assert len(got.junctions) == 0 and len(got.deletions) == 0 and len(got.insertions) == 0
    first, _ = helpers.gpu_segjuncs(P, wl.ref, batches, ctx)
    again, _ = helpers.gpu_segjuncs(P, wl.ref, batches, ctx)        # begin() must reset the sets
    helpers.assert_same_results(first, again, "repeat")
    assert helpers.as_text(first, wl.ref.names)["juncs"] == want["juncs"]
    # idempotence: submitting the same batch twice leaves the sets unchanged (std::set semantics)
    twice, _ = helpers.gpu_segjuncs(P, wl.ref, batches + batches, ctx)
    helpers.assert_same_results(first, twice, "idempotence")
    ctx.close()


def test_sharded_submission_equals_single():
    """Bundle ranges submitted in any order give the same sets (what the multi-GPU path relies on)."""
    wl, P, _ = helpers.load_golden("indel_heavy")
    batches = helpers.pack_both(wl)
    whole, _ = helpers.gpu_segjuncs(P, wl.ref, batches)
    parts = [shard.shard_batch(b, r, 3) for r in (2, 0, 1) for b in batches]
    got, _ = helpers.gpu_segjuncs(P, wl.ref, parts)
    helpers.assert_same_results(got, whole, "3 shards")


def test_device_resident_submit():
    wl, P, _ = helpers.load_golden("wide_flank")
    batches = helpers.pack_both(wl)
    want, _ = pyoracle.segjuncs(P, wl.ref, batches)
    ctx = capi.Context(0); ctx.ref_upload(wl.ref); ctx.segjuncs_begin(P)
    keep = []
    for b in batches:
        t = {k: torch.from_numpy(np.ascontiguousarray(getattr(b, k)).view(np.uint8).reshape(-1)).cuda() for k in
             ("bundles", "seg_count", "reads", "hits", "partner_hits")}
        keep.append(t)
        bc = capi.batch_c(b)
        bc.bundles, bc.seg_count, bc.reads = t["bundles"].data_ptr(), t["seg_count"].data_ptr(), t["reads"].data_ptr()
        bc.hits, bc.partner_hits = t["hits"].data_ptr(), t["partner_hits"].data_ptr()
        torch.cuda.synchronize()
        ctx.segjuncs_submit_device(bc)
    got = ctx.segjuncs_finish()
    helpers.assert_same_results(got, want, "device submit")
    ctx.close()


def test_edge_hits_and_multihit_guard():
    """Windows that touch the contig ends are skipped (segment_juncs.cpp:2154) and reads with more than
    max_seg_multihits hits in a segment are skipped (3499-3506); reads of ragged length share a batch."""
    rng = np.random.default_rng(9)
    ref = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 5000)])
    reads = []
    # hit at the very start / very end of the contig
    reads.append(dict(seq=ref[0:50] + ref[900:951], hits=[[(1, 0, 25, 0, 0)], [(1, 25, 25, 0, 0)], [(1, 900, 25, 0, 0)], [(1, 925, 26, 0, 0)]]))
    reads.append(dict(seq=ref[3000:3050] + ref[4949:5000], hits=[[(1, 3000, 25, 0, 0)], [(1, 3025, 25, 0, 0)], [(1, 4949, 25, 0, 0)], [(1, 4974, 26, 0, 0)]]))
    # 41 hits in one segment
    many = [(1, 100 + 3 * k, 25, 0, 0) for k in range(41)]
    reads.append(dict(seq=ref[1000:1050] + ref[2000:2051], hits=[[(1, 1000, 25, 0, 0)], many, [(1, 2000, 25, 0, 0)], [(1, 2025, 26, 0, 0)]]))
    # short (ragged) read of 60 bases in the same batch, antisense hits
    reads.append(dict(seq=ref[1500:1560], hits=[[(1, 4000, 25, 0, 1)], [(1, 1500, 25, 0, 1)], [], []]))
    # hits on a contig id known from the header only (no sequence)
    reads.append(dict(seq=ref[1000:1050] + ref[2000:2051], hits=[[(2, 1000, 25, 0, 0)], [(2, 1025, 25, 0, 0)], [(2, 2000, 25, 0, 0)], [(2, 2025, 26, 0, 0)]]))
    refimg, batch = helpers.manual_workload([("c1", ref), ("c2", b"")], reads)
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    got, _ = helpers.gpu_segjuncs(P, refimg, [batch])
    want, _ = pyoracle.segjuncs(P, refimg, [batch])
    helpers.assert_same_results(got, want, "edges")


def test_errors_are_loud(lib):
    ctx = capi.Context(0)
    wl, P, _ = helpers.load_golden("splice_2contig")
    b = helpers.pack_both(wl)[0]
    with pytest.raises(capi.ThbError):          # no reference uploaded
        ctx.segjuncs_begin(P); ctx.segjuncs_submit(b)
    ctx.ref_upload(wl.ref)
    with pytest.raises(capi.ThbError):          # outside the GPU path: segment length > 32
        ctx.segjuncs_begin(capi.default_params(segment_length=40))
    ctx.close()


def test_medium_workload_full_pipeline_properties():
    """200k pairs: parity with the oracle plus size-independent properties (sorted unique sets; every junction
    spans [min_segment_intron, max_segment_intron + segment_length + 16])."""
    wl = synth.generate(synth.SynthConfig(contig_lens=(8_000_000, 3_000_000), n_pairs=200_000, seed=11, indel_prob=0.1), workers=4)
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    batches = helpers.pack_both(wl)
    got, t = helpers.gpu_segjuncs(P, wl.ref, batches)
    want, _ = pyoracle.segjuncs(P, wl.ref, batches)
    helpers.assert_same_results(got, want, "200k")
    j = got.junctions
    key = (j["ref_id"].astype(np.int64) << 40) | (j["left"].astype(np.int64) << 8) | 0
    assert (np.diff(key) >= 0).all()
    span = j["right"].astype(np.int64) - j["left"].astype(np.int64)
    assert span.min() >= 50 - 16 and span.max() <= 500000 + 25 + 16


def test_global_coordinates_beyond_32_bits():
    """hg38-scale addressing: a contig placed past global base 5e9 (39-bit global coordinates in keys, window tasks and
    reference fetches).  The image is sparse: two small contigs, the second one far out."""
    wl = synth.generate(synth.SynthConfig(contig_lens=(300_000, 200_000), n_pairs=4000, seed=221, indel_prob=0.3))
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    batches = helpers.pack_both(wl)
    want, _ = pyoracle.segjuncs(P, wl.ref, batches)
    ref = wl.ref
    far = 5_000_000_000 // 64 * 64
    nb2 = (far + ((int(ref.contig_len[1]) + 63) // 64 + 1) * 64) // 64 + 1
    planes = np.zeros(2 * nb2, dtype="<u8"); nmask = np.zeros(nb2, dtype="<u8")
    b0 = int(ref.contig_start[1]) // 64; n1 = ref.n_blocks - b0
    planes[:2 * b0] = ref.planes[:2 * b0]; nmask[:b0] = ref.nmask[:b0]
    planes[2 * (far // 64): 2 * (far // 64) + 2 * n1] = ref.planes[2 * b0:]; nmask[far // 64: far // 64 + n1] = ref.nmask[b0:]
    moved = synth.RefImage(ref.names, ref.contig_len, np.array([0, far], dtype="<u8"), planes, nmask, None)
    got, _ = helpers.gpu_segjuncs(P, moved, batches)
    helpers.assert_same_results(got, want, "far contig")
    assert (got.junctions["ref_id"] == 2).sum() > 20


@pytest.mark.parametrize("name", helpers.reference_input_cases())
def test_gpu_on_the_references_own_test_inputs(name):
    """The reference's own fusion_test/ read sets (BASELINE configs[0] / configs[4]): single-end, --bowtie1, --fusion-search
    --fusion-min-dist 500 --max-intron-length 500; all four outputs equal the reference binary's (tests/golden/reference_*)."""
    wl, P, want = helpers.load_reference_input_case(name)
    batch = synth.pack_side(wl.left, None, False, True)
    got, _ = helpers.gpu_segjuncs(P, wl.ref, [batch])
    txt = helpers.reference_input_texts(got, wl.ref.names)
    for k in want:
        assert txt[k] == want[k], "%s: segment.%s differs from the reference binary's output" % (name, k)
    oracle, _ = pyoracle.segjuncs(P, wl.ref, [batch])
    helpers.assert_same_results(got, oracle, name)
