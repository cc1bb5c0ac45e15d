"""Shared helpers of the parity tests (test infrastructure; may import oracle/)."""
import json
import os

import numpy as np

from tophat_b200 import capi, synth
from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

LIBTYPE = {"fr-unstranded": 1, "fr-firststrand": 2, "fr-secondstrand": 3}


def golden_cases():
    """segment_juncs goldens (outputs of the reference binary); the join goldens are the join_* directories"""
    return sorted(d for d in os.listdir(GOLDEN) if os.path.exists(os.path.join(GOLDEN, d, "segment.juncs")) and
                  not os.path.exists(os.path.join(GOLDEN, d, "inputs.npz")))


def join_golden_cases():
    return sorted(d for d in os.listdir(GOLDEN) if os.path.exists(os.path.join(GOLDEN, d, "left.records.tsv")))


def check_join_golden(name):
    """Both stages through the C ABI on the golden's workload; the joined alignments (after the worker's sort / unique / filters)
    must equal the records the reference's long_spanning_reads wrote (tests/golden/<name>/{left,right}.records.tsv)."""
    from test_gpu_join import joined_to_keys
    d = os.path.join(GOLDEN, name)
    cfg = json.load(open(os.path.join(d, "config.json")))
    kw = dict(cfg["synth"]); kw["contig_lens"] = tuple(kw["contig_lens"])
    wl = synth.generate(synth.SynthConfig(**kw))
    P = capi.default_params(inner_dist_mean=cfg["inner_dist_mean"], inner_dist_std_dev=cfg["inner_dist_std_dev"])
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    res, _ = gpu_segjuncs(P, wl.ref, pack_both(wl), ctx)
    juncs, ins = capi.join_sets_from_results(res)
    ctx.join_begin(P, juncs, ins)
    names = wl.ref.names
    total = 0
    for sname, side in (("left", wl.left), ("right", wl.right)):
        batch = synth.pack_join_side(wl, side, res.junctions)
        got = joined_to_keys(ctx.join_submit(batch), batch, P)
        want = set()
        for line in open(os.path.join(d, sname + ".records.tsv")):
            q, c, pos, cig, flag, nm = line.rstrip("\n").split("\t")
            want.add((int(q), names.index(c) + 1, int(pos), cig, int(flag), int(nm)))
        assert got == want, "%s %s: only ours %r; only reference %r" % (name, sname, sorted(got - want)[:3], sorted(want - got)[:3])
        total += len(want)
    ctx.close()
    assert total > 500


def load_golden(name):
    d = os.path.join(GOLDEN, name)
    cfg = json.load(open(os.path.join(d, "config.json")))
    kw = dict(cfg["synth"])
    kw["contig_lens"] = tuple(kw["contig_lens"])
    wl = synth.generate(synth.SynthConfig(**kw))
    over = dict(inner_dist_mean=cfg["inner_dist_mean"], inner_dist_std_dev=cfg["inner_dist_std_dev"])
    ex = cfg.get("extra", [])
    if "--library-type" in ex:
        over["library_type"] = LIBTYPE[ex[ex.index("--library-type") + 1]]
    if "--fusion-search" in ex:
        over["fusion_search"] = 1
        for opt, field in (("--fusion-anchor-length", "fusion_anchor_length"), ("--fusion-min-dist", "fusion_min_dist")):
            if opt in ex:
                over[field] = int(ex[ex.index(opt) + 1])
    P = capi.default_params(**over)
    texts = {k: open(os.path.join(d, "segment." + k)).read() for k in ("juncs", "insertions", "deletions", "fusions")
             if os.path.exists(os.path.join(d, "segment." + k))}
    return wl, P, texts


def pack_both(wl, fusion_search=False):
    fusion_search = bool(getattr(fusion_search, "fusion_search", fusion_search))      # accepts the Params struct too
    bl = synth.pack_side(wl.left, wl.right, False, fusion_search)
    br = synth.pack_side(wl.right, wl.left, True, fusion_search, order_base=bl.n_bundles)
    return [bl, br]


def as_text(res, names):
    return {"fusions": pyoracle.format_fusions(res.fusions, res.junctions, names),
            "juncs": pyoracle.format_juncs(res.junctions, names),
            "insertions": pyoracle.format_insertions(res.insertions, names),
            "deletions": pyoracle.format_deletions(res.deletions, names)}


def assert_same_results(got, want, what=""):
    for name in ("junctions", "deletions", "insertions", "fusions"):
        a, b = getattr(got, name), getattr(want, name)
        assert a.shape == b.shape, "%s %s: %d vs %d records" % (what, name, len(a), len(b))
        assert (a == b).all(), "%s %s differ" % (what, name)


def gpu_segjuncs(P, ref, batches, ctx=None):
    own = ctx is None
    if own:
        ctx = capi.Context(0)
        ctx.ref_upload(ref)
    ctx.segjuncs_begin(P)
    for b in batches:
        ctx.segjuncs_submit(b)
    got = ctx.segjuncs_finish()
    t = ctx.timing()
    if own:
        ctx.close()
    return got, t


# ---- hand-built bundles for known-answer tests -----------------------------

def manual_workload(contigs, reads):
    """contigs: list of (name, ascii bytes); reads: list of dict(seq=ascii, hits=[[(ref_id,left,len,edit,anti)]*]*nseg,
    partner=[(ref_id,left,len,edit,anti)], flags=int).  Returns (RefImage, PackedBatch)."""
    ref = synth.build_ref_image([n for n, _ in contigs], [synth.codes_from_ascii(s) for _, s in contigs])
    nseg = len(reads[0]["hits"])
    L = max(len(r["seq"]) for r in reads)
    rw = (L + 63) // 64
    bundles = np.zeros(len(reads), dtype=synth.BUNDLE_DTYPE)
    seg_count = np.zeros((len(reads), nseg), dtype="<u2")
    rd = np.zeros((len(reads), 3 * rw), dtype="<u8")
    hits, partner = [], []
    for i, r in enumerate(reads):
        codes = synth.codes_from_ascii(r["seq"])
        padded = np.zeros((1, L), dtype=np.uint8)
        padded[0, :len(codes)] = codes
        rd[i] = synth.pack_reads(padded, rw)[0]
        bundles[i]["read_id"] = i + 1
        bundles[i]["hit_begin"] = len(hits)
        bundles[i]["partner_begin"] = len(partner)
        bundles[i]["read_len"] = len(codes)
        bundles[i]["flags"] = r.get("flags", synth.B_INDELS | synth.B_GAPS)
        for s, hs in enumerate(r["hits"]):
            seg_count[i, s] = len(hs)
            for (rid, left, ln, ed, anti) in hs:
                hits.append((rid, left, left + ln, ln, ed, (synth.HIT_ANTISENSE if anti else 0) | (synth.HIT_END if s == nseg - 1 else 0), 0))
        ph = r.get("partner", [])
        bundles[i]["n_partner"] = len(ph)
        for (rid, left, ln, ed, anti) in ph:
            partner.append((rid, left, left + ln, ln, ed, (synth.HIT_ANTISENSE if anti else 0) | synth.HIT_END, 0))
    H = np.array(hits, dtype=synth.HIT_DTYPE) if hits else np.zeros(0, dtype=synth.HIT_DTYPE)
    PH = np.array(partner, dtype=synth.HIT_DTYPE) if partner else np.zeros(0, dtype=synth.HIT_DTYPE)
    return ref, synth.PackedBatch(nseg, rw, bundles, np.ascontiguousarray(seg_count), np.ascontiguousarray(rd), H, PH, 0)


def our_bin(prog):
    """Path of our drop-in executable; under `pytest --emu` (development aid) the one linked against the host-emulated library."""
    if os.environ.get("THB_TEST_EMU") == "1":
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        return build_emu.build_cli(prog)
    from tophat_b200 import build
    build.build_all()
    return os.path.join(build.BIN_DIR, prog)


# ---- goldens made from the reference's own test inputs (fusion_test/, scripts/make_fusion_test_golden.py) -----------------------

def reference_input_cases():
    return sorted(d for d in os.listdir(GOLDEN) if d.startswith("reference_") and os.path.exists(os.path.join(GOLDEN, d, "inputs.npz")))


def load_reference_input_case(name):
    """-> (Workload with an empty right side, Params of the tophat command line in fusion_test/run_test.sh, expected segment.* texts)"""
    d = os.path.join(GOLDEN, name)
    z = np.load(os.path.join(d, "inputs.npz"))
    lens = [int(x) for x in z["contig_lens"]]
    codes, off = [], 0
    for n in lens:
        codes.append(z["contig_codes"][off:off + n]); off += n
    ref = synth.build_ref_image([str(x) for x in z["contig_names"]], codes)
    reads = z["reads"]; L = reads.shape[1]
    nseg = sum(1 for k in z.files if k.startswith("seg"))
    left = synth.SideData(reads, np.arange(1, reads.shape[0] + 1, dtype="<u4"), [z["seg%d" % k] for k in range(nseg)], z["mapped_hits"], z["unmapped"])
    e = np.zeros(0, dtype=synth.SEGHIT_DTYPE)
    right = synth.SideData(np.zeros((0, L), dtype=np.uint8), np.zeros(0, dtype="<u4"), [e] * nseg, e, np.zeros(0, dtype=bool))
    wl = synth.Workload(synth.SynthConfig(contig_lens=tuple(lens), n_pairs=reads.shape[0], read_len=L, segment_length=25), ref, left, right,
                        np.zeros((0, 4), dtype=np.int64))
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20, bowtie2=0, fusion_search=1, fusion_min_dist=500,
                            max_segment_intron_length=500, max_report_intron_length=500)
    texts = {k: open(os.path.join(d, "segment." + k)).read() for k in ("juncs", "insertions", "deletions", "fusions")}
    return wl, P, texts


def reference_input_texts(res, names):
    t = as_text(res, names)
    t["fusions"] = pyoracle.format_fusions(res.fusions, res.junctions, names, resolve_conflicts=False)    # --fusion-do-not-resolve-conflicts
    return t


# ---- the reference's own fusion_test read sets through BOTH stages (BASELINE configs[0] / configs[4]) ----------------------

FUSION_TEST_EXTRA = ["--bowtie1", "--fusion-search", "--fusion-min-dist", "500", "--fusion-do-not-resolve-conflicts"]
FUSION_TEST_OVERRIDES = {"--max-report-intron": "500", "--max-segment-intron": "500", "--max-coverage-intron": "500", "--max-closure-intron": "500"}


def fusion_test_options():
    o = pyoracle.tophat_common_opts(50, 20, list(FUSION_TEST_EXTRA))
    for k, v in FUSION_TEST_OVERRIDES.items():
        o[o.index(k) + 1] = v
    return o


def _brute_contig_hits(contig_codes, seg, max_mm=2):
    """(contig index, pos0, anti, aligned codes, mismatch mask) of every un-gapped placement with <= max_mm mismatches"""
    out = []
    L = seg.shape[0]
    for anti, q in ((False, seg), (True, synth.revcomp_codes(seg))):
        for ci, codes in enumerate(contig_codes):
            if codes.shape[0] < L:
                continue
            win = np.lib.stride_tricks.sliding_window_view(codes, L)
            mmask = (win != q[None, :]) | (win > 3) | (q[None, :] > 3)
            for pos in np.nonzero(mmask.sum(axis=1) <= max_mm)[0]:
                out.append((ci, int(pos), anti, q, mmask[pos], codes[pos:pos + L]))
    return out


def fusion_test_pipeline(name, td):
    """One of the committed fusion_test input sets (tests/golden/reference_*/inputs.npz) taken through the stage boundary with the
    reference's own tools: prep_reads / fix_map_ordering BAMs, the reference's segment_juncs, its juncs_db over all four outputs (junction,
    insertion, deletion AND fusion contigs), and -- standing in for bowtie against that index -- every un-gapped <= 2-mismatch placement
    of every segment of the unmapped reads on those contigs.  Returns what run_long_spanning_reads needs."""
    import subprocess
    wl, P, _ = load_reference_input_case(name)
    opts = fusion_test_options()
    files = {"fasta": os.path.join(td, "ref.fa"), "header": os.path.join(td, "hdr.sam"), "left_fq": os.path.join(td, "left.fq")}
    synth.write_fasta(files["fasta"], wl.ref); synth.write_sam_header(files["header"], wl.ref); synth.write_fastq(files["left_fq"], wl.left)
    nseg = len(wl.left.seg_hits)
    sams = {"mapped": os.path.join(td, "left_mapped.sam")}
    synth.write_hits_sam(sams["mapped"], wl, wl.left, wl.left.mapped_hits, None)
    for k in range(nseg):
        sams["seg%d" % (k + 1)] = os.path.join(td, "left_seg%d.sam" % (k + 1))
        synth.write_hits_sam(sams["seg%d" % (k + 1)], wl, wl.left, wl.left.seg_hits[k], k)
    bams = {"left_reads": os.path.join(td, "left_kept_reads.bam")}
    subprocess.run([os.path.join(pyoracle.REF_DIR, "prep_reads"), "--sam-header", files["header"], "--outfile", bams["left_reads"],
                    "--index-outfile", bams["left_reads"] + ".index", "--aux-outfile", os.path.join(td, "left.info"), files["left_fq"]],
                   check=True, stderr=subprocess.DEVNULL)
    for key, sam in sams.items():
        bam = os.path.join(td, "left_kept_reads_%s.bam" % key)
        subprocess.run([os.path.join(pyoracle.REF_DIR, "fix_map_ordering"), "--sam-header", files["header"], "--index-outfile", bam + ".index", sam, bam],
                       check=True, stderr=subprocess.DEVNULL)
        bams["left_" + key] = bam
    outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg, opts=opts, paired=False)
    # the junction index: juncs_db <min_anchor> <read length> juncs insertions deletions fusions ref.fa  (tophat.py:2546-2600)
    fa = os.path.join(td, "segment_juncs.fa")
    with open(fa, "w") as f:
        subprocess.run([os.path.join(pyoracle.REF_DIR, "juncs_db"), "3", "26", outs["juncs"], outs["insertions"], outs["deletions"], outs["fusions"],
                        files["fasta"]], check=True, stdout=f, stderr=subprocess.DEVNULL)
    cnames, cseqs = [], []
    for line in open(fa):
        line = line.rstrip("\n")
        if line.startswith(">"):
            cnames.append(line[1:]); cseqs.append("")
        elif cnames:
            cseqs[-1] += line
    ccodes = [synth.codes_from_ascii(s.encode()) for s in cseqs]
    hdr = os.path.join(td, "segment_juncs.hdr.sam")
    with open(hdr, "w") as f:
        f.write("@HD\tVN:1.0\tSO:unsorted\n")
        for n, s in zip(cnames, cseqs):
            f.write("@SQ\tSN:%s\tLN:%d\n" % (n, len(s)))
    offs, lens = synth.segment_layout(wl.cfg.read_len, wl.cfg.segment_length)
    jin = {"juncs_fa": fa, "juncs_header": hdr, "n_contigs": len(cnames), "n_fus_contigs": sum(1 for n in cnames if "|fus|" in n), "left_n_spliced": 0}
    for k in range(nseg):
        sam = os.path.join(td, "left_seg%d.to_spliced.sam" % (k + 1))
        with open(sam, "w") as f:
            for ri in np.nonzero(wl.left.unmapped)[0]:
                seg = wl.left.reads[ri, int(offs[k]):int(offs[k]) + int(lens[k])]
                for (ci, pos, anti, q, mmask, cwin) in _brute_contig_hits(ccodes, seg):
                    md, run = [], 0
                    for x in range(q.shape[0]):
                        if mmask[x]:
                            md.append(str(run)); md.append(chr(synth.CODE2CHAR[min(int(cwin[x]), 4)])); run = 0
                        else:
                            run += 1
                    md.append(str(run)); nm = int(mmask.sum())
                    f.write("%d|%d:%d:%d\t%d\t%s\t%d\t255\t%dM\t*\t0\t0\t%s\t%s\tAS:i:%d\tXN:i:0\tXM:i:%d\tXO:i:0\tXG:i:0\tNM:i:%d\tMD:Z:%s\tYT:Z:UU\n" % (
                        ri + 1, int(offs[k]), k, nseg, 16 if anti else 0, cnames[ci], pos + 1, q.shape[0], synth.CODE2CHAR[q].tobytes().decode(),
                        "I" * q.shape[0], -6 * nm, nm, nm, "".join(md)))
                    jin["left_n_spliced"] += 1
        bam = os.path.join(td, "left_kept_reads_seg%d.to_spliced.bam" % (k + 1))
        subprocess.run([os.path.join(pyoracle.REF_DIR, "fix_map_ordering"), "--sam-header", hdr, "--index-outfile", bam + ".index", sam, bam],
                       check=True, stderr=subprocess.DEVNULL)
        jin["left_spl%d" % (k + 1)] = bam
    return wl, files, bams, jin, outs, nseg, opts
