"""Generates tests/golden/*: outputs of the reference's OWN segment_juncs binary (oracle/_ref, built
from /root/reference/src by oracle/Makefile.ref) on seeded synthetic workloads.

Run in the build container (needs oracle/_ref):   python scripts/make_golden.py
The workloads are re-generated deterministically by the tests from the JSON config stored beside
each output (tophat_b200/synth.py, seeded PCG64), so only the small text outputs are committed.
"""
import json, os, shutil, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tophat_b200 import synth
from oracle import pyoracle

CASES = {
    # name: (SynthConfig kwargs, inner_dist_mean, inner_dist_std_dev, extra options)
    "splice_2contig": (dict(contig_lens=(400_000, 150_000), n_pairs=3000, seed=101), 50, 20, []),
    "indel_heavy": (dict(contig_lens=(300_000,), n_pairs=2500, seed=102, indel_prob=0.5), 50, 20, []),
    "wide_flank": (dict(contig_lens=(500_000,), n_pairs=2000, seed=103, indel_prob=0.1, n_rate=0.004), 200, 20, []),
    "firststrand": (dict(contig_lens=(300_000, 100_000), n_pairs=2000, seed=104), 50, 20, ["--library-type", "fr-firststrand"]),
    # BASELINE configs[4] style: chimeric fragments (ff / fr / rf / rr, intra- and inter-contig) with --fusion-search
    "fusion_inter_intra": (dict(contig_lens=(250_000, 90_000, 60_000), n_pairs=3000, seed=105, indel_prob=0.1, fusion_frac=0.15), 50, 20,
                           ["--fusion-search", "--fusion-anchor-length", "20", "--fusion-min-dist", "20000"]),
    "fusion_default_dist": (dict(contig_lens=(200_000, 120_000), n_pairs=2500, seed=106, fusion_frac=0.25, decoy_rate=1.0), 50, 20,
                            ["--fusion-search"]),
}


def main():
    assert pyoracle.build_reference(), "oracle/_ref missing and /root/reference not available"
    gdir = os.path.join(ROOT, "tests", "golden")
    for name, (kw, im, isd, extra) in CASES.items():
        out = os.path.join(gdir, name)
        os.makedirs(out, exist_ok=True)
        wl = synth.generate(synth.SynthConfig(**kw))
        with tempfile.TemporaryDirectory() as td:
            files = synth.write_pipeline_files(wl, td)
            nseg = len(wl.left.seg_hits)
            bams = pyoracle.make_bams(files, td, nseg)
            outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg,
                                              opts=pyoracle.tophat_common_opts(im, isd, extra))
            for k in ("juncs", "insertions", "deletions") + (("fusions",) if "--fusion-search" in extra else ()):
                shutil.copy(outs[k], os.path.join(out, "segment." + k))
        with open(os.path.join(out, "config.json"), "w") as f:
            json.dump(dict(synth=kw, inner_dist_mean=im, inner_dist_std_dev=isd, extra=extra,
                           generator="scripts/make_golden.py", binary="oracle/_ref/segment_juncs (TopHat 2.1.2, -p1)"), f, indent=1)
        print(name, {k: sum(1 for _ in open(os.path.join(out, "segment." + k))) for k in ("juncs", "insertions", "deletions", "fusions")
                     if os.path.exists(os.path.join(out, "segment." + k))})


JOIN_CASES = {
    # long_spanning_reads: records of the reference binary (qname, contig, pos, cigar, flag, NM) per mate side
    "join_splice_indel": (dict(contig_lens=(200_000, 80_000), n_pairs=1500, seed=111, indel_prob=0.4, keep_truth=True), 50, 20),
}


JOIN_FUSION_CASES = {
    # long_spanning_reads --fusion-search (NOT built on the GPU path yet: DESIGN.md sections 2 / 10): the reference's records incl. the
    # two-part fusion alignments (XF tag) for chimeric reads, committed so that the next round starts from pinned expectations
    "join_fusion_chimeric": (dict(contig_lens=(250_000, 90_000, 60_000), n_pairs=3000, seed=41, indel_prob=0.1, fusion_frac=0.15, keep_truth=True), 50, 20,
                             ["--fusion-search", "--fusion-min-dist", "20000"]),
}


def main_join_fusion():
    gdir = os.path.join(ROOT, "tests", "golden")
    for name, (kw, im, isd, extra) in JOIN_FUSION_CASES.items():
        out = os.path.join(gdir, name)
        os.makedirs(out, exist_ok=True)
        wl = synth.generate(synth.SynthConfig(**kw))
        opts = pyoracle.tophat_common_opts(im, isd, extra)
        with tempfile.TemporaryDirectory() as td:
            files = synth.write_pipeline_files(wl, td)
            nseg = len(wl.left.seg_hits)
            bams = pyoracle.make_bams(files, td, nseg)
            outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg, opts=opts)
            jin = pyoracle.make_join_inputs(wl, files, outs, td, nseg)
            for side in ("left", "right"):
                bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg,
                                                       side=side, tag=".ref", opts=opts, fusions=outs["fusions"])
                _, recs = pyoracle.read_bam(bam)
                with open(os.path.join(out, side + ".fusion_records.tsv"), "w") as f:
                    for r in recs:
                        f.write("%s\t%s\t%d\t%s\t%d\t%d\t%s\n" % (r[0], r[2], r[3], r[5], r[1], r[11]["NM"], r[11].get("XF", "-")))
                print(name, side, len(recs), "records,", sum(1 for r in recs if "XF" in r[11]), "with XF")
        with open(os.path.join(out, "config.json"), "w") as f:
            json.dump(dict(synth=kw, inner_dist_mean=im, inner_dist_std_dev=isd, extra=extra, generator="scripts/make_golden.py",
                           status="expectations for the fusion path of the join, which the GPU path does not implement yet",
                           binary="oracle/_ref/segment_juncs + juncs_db + long_spanning_reads --fusion-search (TopHat 2.1.2, -p1)"), f, indent=1)


def main_join():
    gdir = os.path.join(ROOT, "tests", "golden")
    for name, (kw, im, isd) in JOIN_CASES.items():
        out = os.path.join(gdir, name)
        os.makedirs(out, exist_ok=True)
        wl = synth.generate(synth.SynthConfig(**kw))
        with tempfile.TemporaryDirectory() as td:
            files = synth.write_pipeline_files(wl, td)
            nseg = len(wl.left.seg_hits)
            bams = pyoracle.make_bams(files, td, nseg)
            outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg,
                                              opts=pyoracle.tophat_common_opts(im, isd))
            jin = pyoracle.make_join_inputs(wl, files, outs, td, nseg)
            for side in ("left", "right"):
                bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg,
                                                       side=side, tag=".ref", opts=pyoracle.tophat_common_opts(im, isd))
                _, recs = pyoracle.read_bam(bam)
                with open(os.path.join(out, side + ".records.tsv"), "w") as f:
                    for r in recs:          # file order = the reference's output order
                        f.write("%s\t%s\t%d\t%s\t%d\t%d\n" % (r[0], r[2], r[3], r[5], r[1], r[11]["NM"]))
                print(name, side, len(recs), "records")
        with open(os.path.join(out, "config.json"), "w") as f:
            json.dump(dict(synth=kw, inner_dist_mean=im, inner_dist_std_dev=isd, generator="scripts/make_golden.py",
                           binary="oracle/_ref/segment_juncs + juncs_db + long_spanning_reads (TopHat 2.1.2, -p1)"), f, indent=1)


if __name__ == "__main__":
    main()
    main_join()
    main_join_fusion()
