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
