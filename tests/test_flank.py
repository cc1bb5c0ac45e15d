"""Junction-flank matcher: the step between the two hot binaries -- juncs_db contigs + the segment search
against them -- as one device pass.  CPU part: the oracle is pinned against the reference's own juncs_db (oracle/_ref/juncs_db)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import flank_oracle, pyoracle

C2A = np.frombuffer(b"ACGTN", dtype=np.uint8)


def random_sets(rng, lens, n_j=60, n_d=12, n_i=12, n_f=16, edge=True):
    """Seeded junction / deletion / insertion / fusion sets in the sets' own orders, with records at and beyond the contig ends"""
    nref = len(lens)

    def coords(n):
        r = rng.integers(1, nref + 1, n)
        L = np.asarray(lens)[r - 1]
        left = (rng.random(n) * L).astype(np.int64)
        if edge:
            k = max(1, n // 6)
            left[:k] = rng.integers(0, 30, k)                       # flank clipped at the contig start
            left[k:2 * k] = L[k:2 * k] - rng.integers(1, 30, k)     # .. and at its end
        return r, left, L
    r, left, L = coords(n_j)
    right = np.minimum(left + rng.integers(2, 400, n_j), L + rng.integers(-3, 3, n_j))
    right = np.maximum(right, 0)
    j = np.unique(np.stack([r, left, right, rng.integers(0, 2, n_j)], axis=1), axis=0)
    r, left, L = coords(n_d)
    d = np.unique(np.stack([r, left, np.minimum(left + rng.integers(2, 5, n_d), L), np.zeros(n_d, np.int64)], axis=1), axis=0)
    r, left, L = coords(n_i)
    ins = {}
    for a, b in zip(r, left):
        s = "".join("ACGT"[x] for x in rng.integers(0, 4, int(rng.integers(1, 4))))
        ins.setdefault((int(a), int(b), len(s)), s)                 # Insertion order compares (refid, left, length) only
    ins = [(k[0], k[1], v) for k, v in sorted(ins.items())]
    r1, l1, L1 = coords(n_f)
    r2, l2, L2 = coords(n_f)
    f = np.unique(np.stack([r1, r2, np.minimum(l1, L1 - 1), np.minimum(l2, L2 - 1), rng.integers(7, 11, n_f)], axis=1), axis=0)
    return j, d, ins, f


def random_reference(rng, lens, n_frac=0.01):
    codes = []
    for n in lens:
        c = rng.integers(0, 4, n).astype(np.uint8)
        c[rng.random(n) < n_frac] = 4
        codes.append(c)
    return ["chr%d" % (i + 1) for i in range(len(lens))], codes


def write_set_files(td, names, j, d, ins, f):
    p = {k: os.path.join(td, k) for k in ("juncs", "dels", "ins", "fus")}
    with open(p["juncs"], "w") as fh:
        for r in j:
            fh.write("%s\t%d\t%d\t%s\n" % (names[r[0] - 1], r[1], r[2], "-" if r[3] else "+"))       # segment_juncs.cpp:5041-5046
    with open(p["dels"], "w") as fh:
        for r in d:
            fh.write("%s\t%d\t%d\n" % (names[r[0] - 1], r[1] + 1, r[2]))       # segment_juncs.cpp:5070-5074 writes left+1
    with open(p["ins"], "w") as fh:
        for (r, left, s) in ins:
            fh.write("%s\t%d\t%d\t%s\n" % (names[r - 1], left, left, s))
    with open(p["fus"], "w") as fh:
        for r in f:
            fh.write("%s\t%d\t%s\t%d\t%s\n" % (names[r[0] - 1], r[2], names[r[1] - 1], r[3], {7: "ff", 8: "fr", 9: "rf", 10: "rr"}[int(r[4])]))
    return p


@pytest.mark.parametrize("seed,max_seg_len", [(1, 26), (2, 25), (3, 34), (4, 20)])
def test_contigs_equal_reference_juncs_db(tmp_path, seed, max_seg_len):
    if not os.path.exists(os.path.join(pyoracle.REF_DIR, "juncs_db")):
        pytest.skip("oracle/_ref/juncs_db not built")
    rng = np.random.default_rng(seed)
    lens = [3000, 1200, 500, 90]
    names, codes = random_reference(rng, lens)
    j, d, ins, f = random_sets(rng, lens)
    td = str(tmp_path)
    with open(os.path.join(td, "ref.fa"), "w") as fh:
        for n, c in zip(names, codes):
            fh.write(">%s\n%s\n" % (n, C2A[c].tobytes().decode()))
    p = write_set_files(td, names, j, d, ins, f)
    want = subprocess.run([os.path.join(pyoracle.REF_DIR, "juncs_db"), "3", str(max_seg_len), p["juncs"], p["ins"], p["dels"], p["fus"],
                           os.path.join(td, "ref.fa")], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    cs = flank_oracle.contigs(names, codes, max_seg_len, 3, j, d, ins, f)
    got = flank_oracle.fasta(cs)
    assert len(cs) > 60
    assert {c["kind"] for c in cs} == {0, 1, 2, 3}
    assert got == want


# ----------------------------------------------------------------------------------------------------------------------------
# the device matcher against the oracle (through the C ABI)

def flank_case(seed, max_mm, seg_bounds, n_reads, lens=(3000, 1200, 500, 90), n_j=60, n_d=12, n_i=12, n_f=16, min_anchor=3, flank=None):
    """Reference + sets + reads: segments cut out of contigs (either strand, 0..max_mm+1 substitutions, now and then an N),
    segments cut out of the genome, and random ones."""
    from tophat_b200 import synth
    rng = np.random.default_rng(seed)
    names, codes = random_reference(rng, list(lens))
    j, d, ins, f = random_sets(rng, list(lens), n_j, n_d, n_i, n_f)
    seg_lens = np.diff(seg_bounds)
    max_seg_len = int(flank or seg_lens.max())          # flank: juncs_db's <read_length> when it is not the longest segment
    cs = flank_oracle.contigs(names, codes, max_seg_len, min_anchor, j, d, ins, f)
    L = int(seg_bounds[-1])
    reads = rng.integers(0, 4, (n_reads, L)).astype(np.uint8)
    for r in range(n_reads):
        for k in range(len(seg_lens)):
            s = int(seg_lens[k]); u = rng.random()
            if u < 0.55 and cs:
                c = cs[int(rng.integers(0, len(cs)))]["codes"]
                if len(c) < s:
                    continue
                o = int(rng.integers(0, len(c) - s + 1))
                w = c[o:o + s].copy()
            elif u < 0.7:
                g = codes[int(rng.integers(0, 3))]
                o = int(rng.integers(0, len(g) - s))
                w = g[o:o + s].copy()
            else:
                continue
            w[w > 3] = rng.integers(0, 4, int((w > 3).sum()))
            for _ in range(int(rng.integers(0, max_mm + 2))):
                x = int(rng.integers(0, s)); w[x] = (w[x] + 1 + rng.integers(0, 3)) % 4
            if rng.random() < 0.08:
                w[int(rng.integers(0, s))] = 4
            if rng.random() < 0.5:
                w = flank_oracle._rc(w)
            reads[r, seg_bounds[k]:seg_bounds[k + 1]] = w
    ref = synth.build_ref_image(names, codes)
    return dict(names=names, codes=codes, ref=ref, j=j, d=d, ins=ins, f=f, contigs=cs, reads=reads, max_seg_len=max_seg_len,
                min_seg_len=int(seg_lens.min()), seg_bounds=[int(x) for x in seg_bounds], min_anchor=min_anchor)


def sets_as_records(case):
    from tophat_b200 import synth
    j = np.zeros(len(case["j"]), synth.JUNCTION_DTYPE)
    for n, col in zip(("ref_id", "left", "right", "antisense"), case["j"].T):
        j[n] = col
    d = np.zeros(len(case["d"]), synth.JUNCTION_DTYPE)
    for n, col in zip(("ref_id", "left", "right", "antisense"), case["d"].T):
        d[n] = col
    i = np.zeros(len(case["ins"]), synth.INSERTION_DTYPE)
    for k, (r, left, s) in enumerate(case["ins"]):
        i[k] = (r, left, len(s), s.encode())
    f = np.zeros(len(case["f"]), synth.FUSION_DTYPE)
    for n, col in zip(("ref_id1", "ref_id2", "left", "right", "dir"), case["f"].T):
        f[n] = col
    return j, d, i, f


def check_flank(case, max_mm, max_hits, ref_n_is_mismatch=0):
    from tophat_b200 import capi, synth
    ctx = capi.Context(0)
    ctx.ref_upload(case["ref"])
    P = capi.FlankParams(max_mm, max_hits, case["min_seg_len"], case["max_seg_len"], case["min_anchor"], ref_n_is_mismatch)
    ctx.flank_begin(P, *sets_as_records(case))
    got_c = ctx.flank_contigs()
    want_c = case["contigs"]
    assert len(got_c) == len(want_c)
    for g, w in zip(got_c, want_c):
        assert (int(g["kind"]), int(g["ref_id"]), int(g["ref_id2"]), int(g["left_start"]), int(g["left"]), int(g["aux"]), int(g["length"])) == \
               (w["kind"], w["ref_id"], w["ref_id2"], w["left_start"], w["left"], w["aux"], len(w["codes"])), w["name"]
        assert int(g["right_end"]) == (w["right_end"] & 0xFFFFFFFF)
        if w["kind"] != flank_oracle.KIND_INS:
            assert int(g["right"]) == w["right"]
        else:
            assert g["ins_seq"].decode() == w["ins"]
    rw = (case["reads"].shape[1] + 63) // 64
    hits = ctx.flank_submit(synth.pack_reads(case["reads"], rw), rw, case["seg_bounds"])
    t = ctx.flank_timing()
    got = np.stack([hits[n].astype(np.int64) for n in ("read", "seg", "contig", "pos", "antisense", "mismatches")], axis=1).reshape(-1, 6)
    want = flank_oracle.search([c["codes"] for c in want_c], case["reads"], case["seg_bounds"], max_mm, max_hits, bool(ref_n_is_mismatch))
    assert got.shape == want.shape, "placements: ours %d, oracle %d" % (got.shape[0], want.shape[0])
    assert (got == want).all()
    assert t.n_hits == len(want) and t.n_contigs == len(want_c)
    # the placements as BowtieHits on the genome (thb_flank_spliced_hits) against the restated SplicedBAMHitFactory
    for ma in (8, 3):
        jh = ctx.flank_spliced_hits(ma)
        assert len(jh) == len(hits)
        kinds_kept = set()
        for h, g in zip(got, jh):
            c = want_c[int(h[2])]
            a, b = case["seg_bounds"][int(h[1])], case["seg_bounds"][int(h[1]) + 1]
            q = case["reads"][int(h[0]), a:b]
            if h[4]:
                q = flank_oracle._rc(q)
            cw = c["codes"][int(h[3]):int(h[3]) + (b - a)]
            mism = (q != cw) | (q > 3) | ((cw > 3) if ref_n_is_mismatch else False)
            w = flank_oracle.spliced_hit(c, int(h[3]), int(h[4]), mism, ma, int(h[1]) == len(case["seg_bounds"]) - 2)
            if w is None:
                assert int(g["n_ops"]) == 0, (c["name"], h)
                continue
            kinds_kept.add(c["kind"])
            ops = [(int(x) & 15, int(x) >> 4) for x in g["ops"][:int(g["n_ops"])]]
            assert (int(g["ref_id"]), int(g["left"]), ops, int(g["flags"]), int(g["mismatches"]), int(g["splice_mms"])) == \
                   (w["ref_id"], w["left"], w["ops"], w["flags"], w["mismatches"], w["splice_mms"]), (c["name"], h)
            if c["kind"] == flank_oracle.KIND_FUS:
                assert int(g["ops"][8]) == w["ref_id2"]
        if len(want_c) > 50 and max_mm >= 2:
            assert kinds_kept == {0, 1, 2, 3}, kinds_kept
    # a second batch on the same index (buffers reused), and the empty batch
    hits2 = ctx.flank_submit(synth.pack_reads(case["reads"][::-1].copy(), rw), rw, case["seg_bounds"])
    assert len(hits2) == len(hits)
    assert len(ctx.flank_submit(np.zeros((0, 3 * rw), "<u8"), rw, case["seg_bounds"])) == 0
    ctx.close()
    return len(want), got


FLANK_CASES = {
    # name: (seed, max_mm, seg_bounds, n_reads, max_hits, ref_n_is_mismatch, kwargs)
    "v2_101bp": (11, 2, [0, 25, 50, 75, 101], 120, 40, 0, {}),
    "v2_bowtie2_n": (12, 2, [0, 25, 50, 75, 101], 80, 40, 1, {}),
    "v2_m2_suppression": (13, 2, [0, 25, 50, 75], 100, 2, 0, {}),
    "v0_exact": (14, 0, [0, 25, 50, 76], 80, 40, 0, {}),
    "v2_long_last_segment": (18, 2, [0, 25, 50, 94], 80, 40, 0, dict(flank=26)),
    "v1": (15, 1, [0, 25, 50, 75, 101], 80, 40, 0, {}),
    "v3_two_word_contigs": (16, 3, [0, 30, 60, 94], 60, 40, 0, {}),
    "v3_direct_buckets": (17, 3, [0, 20, 40], 40, 40, 0, dict(lens=(300000, 100000, 5000, 90), n_j=7000, n_d=100, n_i=100, n_f=100)),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FLANK_CASES))
def test_flank_matcher_equals_oracle(name):
    seed, v, bounds, n_reads, max_hits, npol, kw = FLANK_CASES[name]
    case = flank_case(seed, v, np.asarray(bounds), n_reads, **kw)
    n, got = check_flank(case, v, max_hits, npol)
    assert n > 20, n
    if name == "v2_101bp":
        assert set(np.unique(got[:, 5])) == {0, 1, 2} and set(np.unique(got[:, 4])) == {0, 1}


@pytest.mark.gpu
def test_flank_argument_errors():
    from tophat_b200 import capi, synth
    case = flank_case(5, 2, np.asarray([0, 25, 50]), 4)
    ctx = capi.Context(0)
    rw = 1
    with pytest.raises(capi.ThbError):
        ctx.flank_submit(synth.pack_reads(case["reads"], rw), rw, case["seg_bounds"])          # no index
    with pytest.raises(capi.ThbError):
        ctx.flank_begin(capi.FlankParams(2, 40, 25, 25, 3, 0), *sets_as_records(case))          # no reference
    ctx.ref_upload(case["ref"])
    for bad in (capi.FlankParams(4, 40, 25, 25, 3, 0), capi.FlankParams(2, 40, 12, 25, 3, 0), capi.FlankParams(2, 40, 25, 60, 3, 0),
                capi.FlankParams(2, 0, 25, 25, 3, 0), capi.FlankParams(2, 40, 26, 25, 3, 0)):
        with pytest.raises(capi.ThbError):
            ctx.flank_begin(bad, *sets_as_records(case))
    ctx.flank_begin(capi.FlankParams(2, 40, 25, 25, 3, 0), *sets_as_records(case))
    with pytest.raises(capi.ThbError):
        ctx.flank_submit(synth.pack_reads(case["reads"], rw), rw, [0, 24, 50])                  # 24-base segment, seeds built for 25
    with pytest.raises(capi.ThbError):
        ctx.flank_submit(synth.pack_reads(case["reads"], rw), rw, [0, 25, 50, 75])              # past the read words
    # empty sets: an index without contigs answers with no placements
    e = [np.zeros(0, dt) for dt in (synth.JUNCTION_DTYPE, synth.JUNCTION_DTYPE, synth.INSERTION_DTYPE, synth.FUSION_DTYPE)]
    ctx.flank_begin(capi.FlankParams(2, 40, 25, 25, 3, 0), *e)
    assert len(ctx.flank_contigs()) == 0
    assert len(ctx.flank_submit(synth.pack_reads(case["reads"], rw), rw, case["seg_bounds"])) == 0
    ctx.close()


# ----------------------------------------------------------------------------------------------------------------------------
# both stages with the matcher in between, against the reference's own binaries fed the same placements

def flank_pipeline_check(n_pairs=1200, seed=511, indel_prob=0.4, cli_bin=None):
    """stage 1 (ours) -> thb_flank_* -> join (ours)   versus   segment_juncs -> juncs_db -> [the matcher's placements written as bowtie
    would: SAM records on the contig names of the reference's juncs_db] -> fix_map_ordering -> long_spanning_reads, all reference
    binaries.  Checks (1) our contigs are juncs_db's, name for name; (2) the join over thb_flank_spliced_hits gives the records the
    reference's long_spanning_reads writes, i.e. the kernel's placement -> BowtieHit conversion is SplicedBAMHitFactory's."""
    import tempfile
    from tophat_b200 import capi, synth
    import helpers
    from test_gpu_join import joined_to_keys
    wl = synth.generate(synth.SynthConfig(keep_truth=True, contig_lens=(200_000, 80_000), n_pairs=n_pairs, seed=seed, indel_prob=indel_prob))
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    res, _ = helpers.gpu_segjuncs(P, wl.ref, helpers.pack_both(wl), ctx)
    juncs, ins = capi.join_sets_from_results(res)
    names = wl.ref.names
    offs, lens = synth.segment_layout(wl.cfg.read_len, wl.cfg.segment_length)
    bounds = [int(o) for o in offs] + [int(offs[-1] + lens[-1])]
    nseg = len(lens); rw = (wl.cfg.read_len + 63) // 64
    FP = capi.FlankParams(P.segment_mismatches, P.max_seg_multihits, int(lens.min()), int(lens.max()), 3, 1)   # bowtie2 run: N = mismatch
    ctx.flank_begin(FP, res.junctions, res.deletions, res.insertions, res.fusions)
    contigs = ctx.flank_contigs()
    total = 0
    with tempfile.TemporaryDirectory() as td:
        files = synth.write_pipeline_files(wl, td)
        bams = pyoracle.make_bams(files, td, nseg)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg)
        fa = os.path.join(td, "segment_juncs.fa")
        with open(fa, "w") as f:
            subprocess.run([os.path.join(pyoracle.REF_DIR, "juncs_db"), "3", str(int(lens.max())), outs["juncs"], outs["insertions"], outs["deletions"],
                            "/dev/null", files["fasta"]], check=True, stdout=f, stderr=subprocess.DEVNULL)
        cnames, cseqs = [], []
        for line in open(fa):
            if line.startswith(">"):
                cnames.append(line[1:].rstrip("\n"))
            else:
                cseqs.append(line.rstrip("\n"))
        assert len(cnames) == len(contigs) > 100
        kinds = {"GTAG": 0, "del": 1, "ins": 2}
        for nme, sq, c in zip(cnames, cseqs, contigs):
            t = nme.split("|")
            assert (names.index(t[0]) + 1, int(t[1]), int(t[2].split("-")[0]), int(t[3]), kinds[t[4]], len(sq)) == \
                   (int(c["ref_id"]), int(c["left_start"]), int(c["left"]), int(c["right_end"]), int(c["kind"]), int(c["length"])), nme
        hdr = os.path.join(td, "segment_juncs.hdr.sam")
        with open(hdr, "w") as f:
            f.write("@HD\tVN:1.0\tSO:unsorted\n")
            for n, sq in zip(cnames, cseqs):
                f.write("@SQ\tSN:%s\tLN:%d\n" % (n, len(sq)))
        ccodes = [synth.codes_from_ascii(sq.encode()) for sq in cseqs]
        ctx.join_begin(P, juncs, ins)
        for sname, side in (("left", wl.left), ("right", wl.right)):
            idx = np.nonzero(side.unmapped)[0]
            hits = ctx.flank_submit(synth.pack_reads(side.reads[idx], rw), rw, bounds)
            jh = ctx.flank_spliced_hits(P.min_anchor_len)
            assert len(hits) == len(jh) > 100
            jin = {"juncs_fa": fa, "juncs_header": hdr, "n_contigs": len(cnames)}
            spliced = []
            for k in range(nseg):
                m = hits["seg"] == k
                hk, jk = hits[m], jh[m]
                sam = os.path.join(td, "%s_seg%d.to_spliced.sam" % (sname, k + 1))
                with open(sam, "w") as f:
                    for h in hk:
                        ri = int(idx[int(h["read"])]); s = int(lens[k])
                        q = side.reads[ri, bounds[k]:bounds[k + 1]]
                        if h["antisense"]:
                            q = flank_oracle._rc(q)
                        cw = ccodes[int(h["contig"])][int(h["pos"]):int(h["pos"]) + s]
                        mm = (q != cw) | (q > 3) | (cw > 3)
                        md, run = [], 0
                        for x in range(s):
                            if mm[x]:
                                md.append(str(run)); md.append(chr(synth.CODE2CHAR[min(int(cw[x]), 4)])); run = 0
                            else:
                                run += 1
                        md.append(str(run)); nm = int(mm.sum())
                        assert nm == int(h["mismatches"])
                        f.write("%d|%d:%d:%d\t%d\t%s\t%d\t255\t%dM\t*\t0\t0\t%s\t%s\tAS:i:%d\tXN:i:0\tXM:i:%d\tXO:i:0\tXG:i:0\tNM:i:%d\tMD:Z:%s\tYT:Z:UU\n" % (
                            int(side.ids[ri]), bounds[k], k, nseg, 16 if h["antisense"] else 0, cnames[int(h["contig"])], int(h["pos"]) + 1, s,
                            synth.CODE2CHAR[q].tobytes().decode(), "I" * s, -6 * nm, nm, nm, "".join(md)))
                bam = os.path.join(td, "%s_kept_reads_seg%d.to_spliced.bam" % (sname, k + 1))
                subprocess.run([os.path.join(pyoracle.REF_DIR, "fix_map_ordering"), "--sam-header", hdr, "--index-outfile", bam + ".index", sam, bam],
                               check=True, stderr=subprocess.DEVNULL)
                jin["%s_spl%d" % (sname, k + 1)] = bam
                keep = jk["n_ops"] > 0
                spliced.append((idx[hk["read"][keep]], jk[keep]))
            jin["%s_n_spliced" % sname] = int(len(hits))
            batch = synth.assemble_join_batch(side, spliced)
            got = joined_to_keys(ctx.join_submit(batch), batch, P)
            ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg,
                                                       side=sname, tag=".ref")
            _, recs = pyoracle.read_bam(ref_bam)
            want = set((int(r[0]), names.index(r[2]) + 1, r[3], r[5], r[1], r[11]["NM"]) for r in recs)
            assert got == want, "%s: only ours %r; only reference %r" % (sname, sorted(got - want)[:3], sorted(want - got)[:3])
            if cli_bin:
                # our executable searching the junction index itself (TOPHAT_GPU_FLANK_SEARCH=1, no spliced segment files), -p1 and -p3
                for thr in (1, 3):
                    our_bam = pyoracle.run_long_spanning_reads(cli_bin, files, bams, jin, outs, td, nseg, side=sname, tag=".flank%d" % thr, with_spliced=False,
                                                               threads=thr, env=dict(os.environ, TOPHAT_GPU_FLANK_SEARCH="1"))
                    _, ours = pyoracle.read_bam(our_bam)
                    assert len(ours) == len(recs), "%s -p%d: %d records vs %d in the reference output" % (sname, thr, len(ours), len(recs))
                    for x, y in zip(ours, recs):
                        assert x == y, "record differs:\n ours %r\n ref  %r" % (x, y)
            spl = [r for r in want if "N" in r[3] or "D" in r[3] or "I" in r[3]]
            assert len(spl) > 50, len(spl)
            total += len(want)
    ctx.close()
    return total


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
def test_flank_pipeline_matches_reference_binaries():
    assert flank_pipeline_check() > 500


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
def test_cli_long_spanning_reads_searches_the_junction_index_itself():
    """long_spanning_reads with TOPHAT_GPU_FLANK_SEARCH=1 and no spliced segment files: the records of the reference's binary that was
    given the placements as to_spliced.bam files, in the same order."""
    import helpers
    assert flank_pipeline_check(cli_bin=helpers.our_bin("long_spanning_reads")) > 500


# ----------------------------------------------------------------------------------------------------------------------------
# reads of two lengths in one run (101 bp = 4 segments with a 26-base last one, 75 bp = 3 segments)

def mixed_length_cli_check(cli_bin, n_pairs=600):
    """One reads file with 101-bp and 75-bp reads: the reference chain (segment_juncs, juncs_db 3 26, long_spanning_reads fed the
    placements as BAM files) against our long_spanning_reads searching the index itself -- per read length its own segment layout,
    segment count and end-of-read flag."""
    import tempfile
    from tophat_b200 import capi, synth
    base = dict(contig_lens=(200_000, 80_000), keep_truth=True, indel_prob=0.3)
    cfgs = [synth.SynthConfig(n_pairs=n_pairs, seed=611, read_len=101, **base), synth.SynthConfig(n_pairs=n_pairs, seed=611, read_len=75, **base)]
    rd = synth.make_reference(cfgs[0])
    wls = [synth.generate(c, refdata=rd) for c in cfgs]
    names = wls[0].ref.names
    NSEG = 4
    with tempfile.TemporaryDirectory() as td:
        parts = []
        for k, wl in enumerate(wls):
            d = os.path.join(td, "part%d" % k)
            parts.append(synth.write_pipeline_files(wl, d))
        files = {"fasta": parts[0]["fasta"], "header": parts[0]["header"]}
        offset = [0, wls[0].left.reads.shape[0]]
        for side in ("left", "right"):
            files[side + "_fq"] = os.path.join(td, side + ".fq")
            with open(files[side + "_fq"], "wb") as out:
                for p in parts:
                    out.write(open(p[side + "_fq"], "rb").read())
            for key in ["mapped"] + ["seg%d" % (s + 1) for s in range(NSEG)]:
                fk = "%s_%s_sam" % (side, key)
                files[fk] = os.path.join(td, "%s_%s.sam" % (side, key))
                with open(files[fk], "w") as out:
                    for k, p in enumerate(parts):
                        if fk not in p:
                            continue                                     # the 75-bp reads have no fourth segment
                        for line in open(p[fk]):
                            q, rest = line.split("\t", 1)
                            head, sep, tail = q.partition("|")
                            out.write("%d%s%s\t%s" % (int(head) + offset[k], sep, tail, rest))
        bams = pyoracle.make_bams(files, td, NSEG)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, NSEG)
        fa = os.path.join(td, "segment_juncs.fa")
        with open(fa, "w") as f:
            subprocess.run([os.path.join(pyoracle.REF_DIR, "juncs_db"), "3", "26", outs["juncs"], outs["insertions"], outs["deletions"], "/dev/null",
                            files["fasta"]], check=True, stdout=f, stderr=subprocess.DEVNULL)
        cnames, cseqs = [], []
        for line in open(fa):
            (cnames if line.startswith(">") else cseqs).append(line[1:].rstrip("\n") if line.startswith(">") else line.rstrip("\n"))
        hdr = os.path.join(td, "segment_juncs.hdr.sam")
        with open(hdr, "w") as f:
            f.write("@HD\tVN:1.0\tSO:unsorted\n")
            for n, sq in zip(cnames, cseqs):
                f.write("@SQ\tSN:%s\tLN:%d\n" % (n, len(sq)))
        ccodes = [synth.codes_from_ascii(sq.encode()) for sq in cseqs]
        # the sets as the reference wrote them -> the index (flank 26, bowtie2's N policy as the executable will use)
        P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
        jn = pyoracle.parse_juncs(outs["juncs"], names)
        dl = np.zeros(0, synth.JUNCTION_DTYPE); ins = np.zeros(0, synth.INSERTION_DTYPE)
        dlines = [l.split("\t") for l in open(outs["deletions"]).read().splitlines()]
        if dlines:
            dl = np.zeros(len(dlines), synth.JUNCTION_DTYPE)
            for i, t in enumerate(dlines):
                dl[i] = (names.index(t[0]) + 1, int(t[1]) - 1, int(t[2]), 0)
        ilines = [l.split("\t") for l in open(outs["insertions"]).read().splitlines()]
        if ilines:
            ins = np.zeros(len(ilines), synth.INSERTION_DTYPE)
            for i, t in enumerate(ilines):
                ins[i] = (names.index(t[0]) + 1, int(t[1]), len(t[3]), t[3].encode())
        ctx = capi.Context(0); ctx.ref_upload(wls[0].ref)
        ctx.flank_begin(capi.FlankParams(2, 40, 25, 26, 3, 1), jn, dl, ins, np.zeros(0, synth.FUSION_DTYPE))
        assert len(ctx.flank_contigs()) == len(cnames) > 100
        total = 0
        for sname in ("left", "right"):
            sams = [open(os.path.join(td, "%s_seg%d.to_spliced.sam" % (sname, s + 1)), "w") for s in range(NSEG)]
            n_hits = 0
            for k, wl in enumerate(wls):
                side = wl.left if sname == "left" else wl.right
                offs, lens = synth.segment_layout(wl.cfg.read_len, wl.cfg.segment_length)
                bounds = [int(o) for o in offs] + [int(offs[-1] + lens[-1])]
                rw = (wl.cfg.read_len + 63) // 64
                idx = np.nonzero(side.unmapped)[0]
                hits = ctx.flank_submit(synth.pack_reads(side.reads[idx], rw), rw, bounds)
                n_hits += len(hits)
                for h in hits:
                    s = int(h["seg"]); ri = int(idx[int(h["read"])]); ln = int(lens[s])
                    q = side.reads[ri, bounds[s]:bounds[s + 1]]
                    if h["antisense"]:
                        q = flank_oracle._rc(q)
                    cw = ccodes[int(h["contig"])][int(h["pos"]):int(h["pos"]) + ln]
                    mm = (q != cw) | (q > 3) | (cw > 3)
                    md, run = [], 0
                    for x in range(ln):
                        if mm[x]:
                            md.append(str(run)); md.append(chr(synth.CODE2CHAR[min(int(cw[x]), 4)])); run = 0
                        else:
                            run += 1
                    md.append(str(run)); nm = int(mm.sum())
                    sams[s].write("%d|%d:%d:%d\t%d\t%s\t%d\t255\t%dM\t*\t0\t0\t%s\t%s\tAS:i:%d\tXN:i:0\tXM:i:%d\tXO:i:0\tXG:i:0\tNM:i:%d\tMD:Z:%s\tYT:Z:UU\n" % (
                        ri + 1 + offset[k], bounds[s], s, len(lens), 16 if h["antisense"] else 0, cnames[int(h["contig"])], int(h["pos"]) + 1, ln,
                        synth.CODE2CHAR[q].tobytes().decode(), "I" * ln, -6 * nm, nm, nm, "".join(md)))
            jin = {"juncs_fa": fa, "juncs_header": hdr, "n_contigs": len(cnames)}
            for s, f in enumerate(sams):
                f.close()
                bam = os.path.join(td, "%s_kept_reads_seg%d.to_spliced.bam" % (sname, s + 1))
                subprocess.run([os.path.join(pyoracle.REF_DIR, "fix_map_ordering"), "--sam-header", hdr, "--index-outfile", bam + ".index", f.name, bam],
                               check=True, stderr=subprocess.DEVNULL)
                jin["%s_spl%d" % (sname, s + 1)] = bam
            ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, NSEG, side=sname, tag=".ref")
            _, want = pyoracle.read_bam(ref_bam)
            lens_seen = {len(r[6]) for r in want}
            assert lens_seen == {101, 75}, lens_seen
            assert n_hits > 200 and sum(1 for r in want if "N" in r[5]) > 60
            for thr in (1, 2):
                our_bam = pyoracle.run_long_spanning_reads(cli_bin, files, bams, jin, outs, td, NSEG, side=sname, tag=".flank%d" % thr, with_spliced=False, threads=thr,
                                                           env=dict(os.environ, TOPHAT_GPU_FLANK_SEARCH="1", TOPHAT_GPU_FLANK_LENGTH="26"))
                _, got = pyoracle.read_bam(our_bam)
                assert len(got) == len(want), "%s -p%d: %d records vs %d" % (sname, thr, len(got), len(want))
                for x, y in zip(got, want):
                    assert x == y, "record differs:\n ours %r\n ref  %r" % (x, y)
            total += len(want)
        ctx.close()
    return total


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
def test_cli_in_process_junction_index_with_two_read_lengths():
    import helpers
    assert mixed_length_cli_check(helpers.our_bin("long_spanning_reads")) > 400


# ----------------------------------------------------------------------------------------------------------------------------
# committed golden vectors (scripts/make_flank_golden.py: FASTA of the reference's juncs_db, placements by exhaustion over it)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLANK_GOLDENS = sorted(d for d in os.listdir(GOLDEN_DIR) if d.startswith("flank_"))


def _load_flank_golden(name):
    import json
    d = os.path.join(GOLDEN_DIR, name)
    cfg = json.load(open(os.path.join(d, "config.json")))
    case = flank_case(cfg["seed"], cfg["max_mismatches"], np.asarray(cfg["seg_bounds"]), cfg["n_reads"], **cfg["kwargs"])
    fasta = open(os.path.join(d, "juncs_db.fa")).read()
    want = np.loadtxt(os.path.join(d, "placements.tsv"), dtype=np.int64, ndmin=2).reshape(-1, 6)
    return cfg, case, fasta, want


@pytest.mark.parametrize("name", FLANK_GOLDENS)
def test_flank_oracle_equals_golden(name):
    """no reference binary needed: the restated juncs_db and the exhaustive search reproduce the committed outputs"""
    cfg, case, fasta, want = _load_flank_golden(name)
    assert flank_oracle.fasta(case["contigs"]) == fasta
    got = flank_oracle.search([c["codes"] for c in case["contigs"]], case["reads"], case["seg_bounds"], cfg["max_mismatches"], cfg["max_multihits"],
                              bool(cfg["ref_n_is_mismatch"]))
    assert got.shape == want.shape and (got == want).all()


def flank_golden_check(name):
    from tophat_b200 import capi, synth
    cfg, case, fasta, want = _load_flank_golden(name)
    ctx = capi.Context(0); ctx.ref_upload(case["ref"])
    ctx.flank_begin(capi.FlankParams(cfg["max_mismatches"], cfg["max_multihits"], case["min_seg_len"], case["max_seg_len"], case["min_anchor"], cfg["ref_n_is_mismatch"]),
                    *sets_as_records(case))
    names = [l[1:] for l in fasta.splitlines() if l.startswith(">")]
    seqs = [l for l in fasta.splitlines() if not l.startswith(">")]
    contigs = ctx.flank_contigs()
    assert len(contigs) == len(names)
    for c, nme, sq in zip(contigs, names, seqs):
        t = nme.split("|")
        assert int(c["left_start"]) == int(t[1]) and int(c["length"]) == len(sq), nme
    rw = (case["reads"].shape[1] + 63) // 64
    hits = ctx.flank_submit(synth.pack_reads(case["reads"], rw), rw, case["seg_bounds"])
    got = np.stack([hits[n].astype(np.int64) for n in ("read", "seg", "contig", "pos", "antisense", "mismatches")], axis=1).reshape(-1, 6)
    assert got.shape == want.shape and (got == want).all()
    ctx.close()
    return len(want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FLANK_GOLDENS)
def test_flank_matcher_equals_golden(name):
    assert flank_golden_check(name) > 40
