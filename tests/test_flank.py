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

def flank_case(seed, max_mm, seg_bounds, n_reads, lens=(3000, 1200, 500, 90), n_j=60, n_d=12, n_i=12, n_f=16, min_anchor=3):
    """Reference + sets + reads: segments cut out of contigs (either strand, 0..max_mm+1 substitutions, now and then an N),
    segments cut out of the genome, and random ones."""
    from tophat_b200 import synth
    rng = np.random.default_rng(seed)
    names, codes = random_reference(rng, list(lens))
    j, d, ins, f = random_sets(rng, list(lens), n_j, n_d, n_i, n_f)
    seg_lens = np.diff(seg_bounds)
    max_seg_len = int(seg_lens.max())
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
    "v1": (15, 1, [0, 25, 50, 75, 101], 80, 40, 0, {}),
    "v3_two_word_contigs": (16, 3, [0, 30, 60, 94], 60, 40, 0, {}),
    "v3_direct_buckets": (17, 3, [0, 20, 40], 40, 40, 0, dict(lens=(60000, 20000, 5000, 90), n_j=1500, n_d=100, n_i=100, n_f=100)),
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
        ctx.flank_submit(synth.pack_reads(case["reads"], rw), rw, [0, 25, 51])                  # 26-base segment, index built for 25
    with pytest.raises(capi.ThbError):
        ctx.flank_submit(synth.pack_reads(case["reads"], rw), rw, [0, 25, 50, 75])              # past the read words
    # empty sets: an index without contigs answers with no placements
    e = [np.zeros(0, dt) for dt in (synth.JUNCTION_DTYPE, synth.JUNCTION_DTYPE, synth.INSERTION_DTYPE, synth.FUSION_DTYPE)]
    ctx.flank_begin(capi.FlankParams(2, 40, 25, 25, 3, 0), *e)
    assert len(ctx.flank_contigs()) == 0
    assert len(ctx.flank_submit(synth.pack_reads(case["reads"], rw), rw, case["seg_bounds"])) == 0
    ctx.close()
