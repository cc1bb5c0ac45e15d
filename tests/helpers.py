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
