"""Drop-in boundary test of long_spanning_reads: our executable (C++ host + chain-join kernel) against the reference's
own binary on identical inputs -- segment BAMs, junction-index segment BAMs built from the reference's juncs_db output,
and the segment.juncs/insertions/deletions of segment_juncs.  BAM byte identity is not meaningful (zlib); the decoded
records must be identical, in the same order."""
import os
import tempfile

import pytest

import helpers
from tophat_b200 import build, synth
from oracle import pyoracle

OUR_BIN = os.path.join(build.BIN_DIR, "long_spanning_reads")


def test_host_binary_builds():
    build.build_library()
    outs = build.build_host_binaries()
    assert OUR_BIN in outs and os.access(OUR_BIN, os.X_OK)


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("kw,spliced", [
    (dict(contig_lens=(300_000, 120_000), n_pairs=3000, seed=401), True),
    (dict(contig_lens=(300_000,), n_pairs=3000, seed=402, indel_prob=0.5), True),
    (dict(contig_lens=(250_000,), n_pairs=2500, seed=403, indel_prob=0.2, decoy_rate=2.0, sub_rate=0.01), True),
    (dict(contig_lens=(250_000,), n_pairs=2000, seed=404), False),
    (dict(contig_lens=(250_000,), n_pairs=2000, seed=405, read_len=75, indel_prob=0.3), True),
    (dict(contig_lens=(250_000,), n_pairs=1500, seed=406, read_len=150, indel_prob=0.3, n_rate=0.004), True),
])
def test_cli_matches_reference_binary(kw, spliced):
    OUR_BIN = helpers.our_bin("long_spanning_reads")
    with tempfile.TemporaryDirectory() as td:
        wl = synth.generate(synth.SynthConfig(keep_truth=True, **kw))
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg)
        jin = pyoracle.make_join_inputs(wl, files, outs, td, nseg, max_seg_len=int(max(synth.segment_layout(wl.cfg.read_len, wl.cfg.segment_length)[1])))
        total = 0
        for side in ("left", "right"):
            ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg,
                                                       side=side, tag=".ref", with_spliced=spliced)
            our_bam = pyoracle.run_long_spanning_reads(OUR_BIN, files, bams, jin, outs, td, nseg, side=side, tag=".b200", with_spliced=spliced)
            refs_a, a = pyoracle.read_bam(our_bam)
            refs_b, b = pyoracle.read_bam(ref_bam)
            assert refs_a == refs_b
            assert len(a) == len(b), "%s: %d records vs %d in the reference output" % (side, len(a), len(b))
            for x, y in zip(a, b):
                assert x == y, "record differs:\n ours %r\n ref  %r" % (x, y)
            total += len(b)
        assert total > 200
        if spliced:
            assert jin["left_n_spliced"] + jin["right_n_spliced"] > 100


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
def test_cli_read_id_ranges_do_not_change_the_output():
    """-p N: our executable splits the read ids into N ranges at entries of the reads file's .index (every stream positioned by its
    own index and filtered by id) and writes <out>0.bam .. <out>N-1.bam like the reference (long_spanning_reads.cpp:3056-3064).  The
    concatenation must equal the -p1 output = the reference's -p1 records, whatever N; every side-index entry of every file must be the
    BGZF virtual offset of the first record of that read id."""
    import struct, zlib
    OUR_BIN = helpers.our_bin("long_spanning_reads")
    with tempfile.TemporaryDirectory() as td:
        wl = synth.generate(synth.SynthConfig(keep_truth=True, contig_lens=(400_000, 150_000), n_pairs=9000, seed=407, indel_prob=0.3))
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg)
        jin = pyoracle.make_join_inputs(wl, files, outs, td, nseg)
        ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg, side="left", tag=".ref")
        _, want = pyoracle.read_bam(ref_bam)
        n_files = {}
        for n in (1, 3, 7):
            bam = pyoracle.run_long_spanning_reads(OUR_BIN, files, bams, jin, outs, td, nseg, side="left", tag=".p%d" % n, threads=n)
            _, got = pyoracle.read_bam(bam)
            assert got == want, "-p%d: records differ from the reference" % n
            parts = [bam] if os.path.exists(bam) else []
            while not os.path.exists(bam) and os.path.exists(bam[:-4] + "%d.bam" % len(parts)):
                parts.append(bam[:-4] + "%d.bam" % len(parts))
            n_files[n] = len(parts)
            assert not any(f.endswith(".thb_tmp") for f in os.listdir(td)), "temporary output left behind"
            for part in parts:
                raw = open(part, "rb").read()
                for line in open(part + ".index"):
                    rid, voff = (int(x) for x in line.split("\t"))
                    boff, inoff = voff >> 16, voff & 0xffff
                    data = b""
                    while len(data) < inoff + 64 and boff < len(raw):
                        bsize = struct.unpack_from("<H", raw, boff + 16)[0] + 1
                        data += zlib.decompress(raw[boff + 18: boff + bsize - 8], -15); boff += bsize
                    l_qn = data[inoff + 12]
                    assert int(data[inoff + 36: inoff + 36 + l_qn - 1].decode()) == rid
        assert len(want) > 3000 and n_files[1] == 1 and n_files[3] == 3 and n_files[7] >= 4


def _compare_bams(our_bam, ref_bam, what):
    refs_a, a = pyoracle.read_bam(our_bam)
    refs_b, b = pyoracle.read_bam(ref_bam)
    assert refs_a == refs_b
    assert len(a) == len(b), "%s: %d records vs %d in the reference output" % (what, len(a), len(b))
    for x, y in zip(a, b):
        assert x == y, "%s: record differs:\n ours %r\n ref  %r" % (what, x, y)
    return b


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("threads", [1, 3])
def test_cli_fusion_search_matches_reference_binary(threads):
    """--fusion-search (a7): chimeric synthetic fragments (ff / fr / rf / rr, intra- and inter-contig).  The join closes the fusion
    point against segment.fusions (merge_chain's fusion closure, long_spanning_reads.cpp:1596-1819) and writes every fusion alignment
    as two records with XF tags (bwt_map.cpp:2047-2083); every record -- fields, order, all aux tags -- equals the reference's, and
    equals the expectations pinned in tests/golden/join_fusion_chimeric."""
    import json
    OUR_BIN = helpers.our_bin("long_spanning_reads")
    gdir = os.path.join(helpers.GOLDEN, "join_fusion_chimeric")
    cfg = json.load(open(os.path.join(gdir, "config.json")))
    kw = dict(cfg["synth"]); kw["contig_lens"] = tuple(kw["contig_lens"])
    opts = pyoracle.tophat_common_opts(cfg["inner_dist_mean"], cfg["inner_dist_std_dev"], cfg["extra"])
    with tempfile.TemporaryDirectory() as td:
        wl = synth.generate(synth.SynthConfig(**kw))
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg, opts=opts)
        jin = pyoracle.make_join_inputs(wl, files, outs, td, nseg)
        n_xf = 0
        for side in ("left", "right"):
            ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg,
                                                       side=side, tag=".ref", opts=opts, fusions=outs["fusions"])
            our_bam = pyoracle.run_long_spanning_reads(OUR_BIN, files, bams, jin, outs, td, nseg, side=side, tag=".b200p%d" % threads, opts=opts,
                                                       fusions=outs["fusions"], threads=threads)
            recs = _compare_bams(our_bam, ref_bam, side)
            n_xf += sum(1 for r in recs if "XF" in r[11])
            want = [l.rstrip("\n") for l in open(os.path.join(gdir, side + ".fusion_records.tsv"))]
            got = ["%s\t%s\t%d\t%s\t%d\t%d\t%s" % (r[0], r[2], r[3], r[5], r[1], r[11]["NM"], r[11].get("XF", "-")) for r in pyoracle.read_bam(our_bam)[1]]
            assert got == want, "%s: records differ from the pinned expectations" % side
        assert n_xf > 100


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("name", helpers.reference_input_cases())
def test_cli_fusion_test_sets_through_both_stages(name):
    """BASELINE configs[0] / configs[4]: the reference's own fusion_test read sets (junction / indel / fusion / total, intra- and
    inter-contig) through BOTH stages with the options tophat.py derives from fusion_test/run_test.sh: our segment_juncs reproduces the
    four segment.* files, and our long_spanning_reads -- fed the junction index the reference's juncs_db builds from them, fusion
    contigs included -- reproduces every record of the reference's candidates BAM (two-record XF form for the fusion alignments)."""
    SJ, LSR = helpers.our_bin("segment_juncs"), helpers.our_bin("long_spanning_reads")
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as td:
        wl, files, bams, jin, outs, nseg, opts = helpers.fusion_test_pipeline(name, td)
        ours1 = pyoracle.run_segment_juncs(SJ, files, bams, td, nseg, opts=opts, paired=False, tag=".b200")
        for k in ("juncs", "insertions", "deletions", "fusions"):
            assert open(ours1[k]).read() == open(outs[k]).read(), "segment.%s differs from the reference's" % k
        ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg, side="left",
                                                   tag=".ref", opts=opts, fusions=outs["fusions"])
        our_bam = pyoracle.run_long_spanning_reads(LSR, files, bams, jin, ours1, td, nseg, side="left", tag=".b200", opts=opts, fusions=ours1["fusions"])
        recs = _compare_bams(our_bam, ref_bam, name)
        assert len(recs) > 200 and jin["left_n_spliced"] > 100
        # the same without any junction-index BAM: long_spanning_reads builds and searches the index itself (fusion contigs included)
        flank_bam = pyoracle.run_long_spanning_reads(LSR, files, bams, jin, ours1, td, nseg, side="left", tag=".b200flank", opts=opts, fusions=ours1["fusions"],
                                                     with_spliced=False, env=dict(os.environ, TOPHAT_GPU_FLANK_SEARCH="1", TOPHAT_GPU_FLANK_LENGTH="26"))     # the pipeline above ran juncs_db 3 26
        _compare_bams(flank_bam, ref_bam, name + " (in-process junction index)")
        if "fusion" in name or "total" in name:
            assert jin["n_fus_contigs"] > 10 and sum(1 for r in recs if "XF" in r[11]) > 500


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("gz", [False, True])
def test_cli_reads_as_fastq_text(gz):
    """The reads argument may be FASTQ text (plain or gzip-ed) instead of the BAM prep_reads writes (ReadStream::init, reads.cpp:528-560):
    same records as the reference's binary given the same file."""
    import gzip
    OUR_BIN = helpers.our_bin("long_spanning_reads")
    with tempfile.TemporaryDirectory() as td:
        wl = synth.generate(synth.SynthConfig(keep_truth=True, contig_lens=(250_000,), n_pairs=1500, seed=407, indel_prob=0.3))
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg)
        jin = pyoracle.make_join_inputs(wl, files, outs, td, nseg)
        # the kept reads as text: numeric names, as prep_reads numbers them
        fq = os.path.join(td, "left_kept_reads.fq" + (".gz" if gz else ""))
        with (gzip.open(fq, "wb") if gz else open(fq, "wb")) as f:
            for i in range(wl.left.reads.shape[0]):
                f.write(b"@%d some comment\n" % int(wl.left.ids[i])); f.write(synth.CODE2CHAR[wl.left.reads[i]].tobytes())
                f.write(b"\n+\n"); f.write(bytes(33 + (i + k) % 40 for k in range(wl.left.reads.shape[1]))); f.write(b"\n")
        tb = dict(bams); tb["left_reads"] = fq
        if gz:
            our_bam = pyoracle.run_long_spanning_reads(OUR_BIN, files, tb, jin, outs, td, nseg, side="left", tag=".b200gz")
            plain = dict(bams); plain["left_reads"] = fq[:-3]
            with open(plain["left_reads"], "wb") as f:
                f.write(gzip.open(fq, "rb").read())
            ref_bam = pyoracle.run_long_spanning_reads(OUR_BIN, files, plain, jin, outs, td, nseg, side="left", tag=".b200plain")
        else:
            ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, tb, jin, outs, td, nseg, side="left", tag=".ref")
            our_bam = pyoracle.run_long_spanning_reads(OUR_BIN, files, tb, jin, outs, td, nseg, side="left", tag=".b200")
        _, a = pyoracle.read_bam(our_bam)
        _, b = pyoracle.read_bam(ref_bam)
        assert len(a) == len(b) > 100
        for x, y in zip(a, b):
            assert x == y, "record differs:\n ours %r\n ref  %r" % (x, y)
