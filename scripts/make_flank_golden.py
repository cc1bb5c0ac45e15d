"""Golden vectors for the junction index step (tests/golden/flank_*): the FASTA the reference's own juncs_db writes for seeded sets
(oracle/_ref/juncs_db, built from /root/reference/src/juncs_db.cpp) and the placements of seeded reads on those contigs by exhaustive
search over that FASTA (bowtie's -v/-k/-m contract).  Run where oracle/_ref exists; the outputs are committed.
usage: python scripts/make_flank_golden.py"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from oracle import flank_oracle, pyoracle  # noqa: E402
import test_flank  # noqa: E402

CASES = {"flank_v2_101bp": ("v2_101bp", 2, 40, 0), "flank_v3_two_word": ("v3_two_word_contigs", 3, 40, 0), "flank_v2_m2": ("v2_m2_suppression", 2, 2, 0)}
C2A = np.frombuffer(b"ACGTN", dtype=np.uint8)


def main():
    import tempfile
    for gname, (cname, v, k, npol) in CASES.items():
        seed, v_, bounds, n_reads, max_hits, npol_, kw = test_flank.FLANK_CASES[cname]
        case = test_flank.flank_case(seed, v_, np.asarray(bounds), n_reads, **kw)
        d = os.path.join(ROOT, "tests", "golden", gname)
        os.makedirs(d, exist_ok=True)
        with tempfile.TemporaryDirectory() as td:
            with open(os.path.join(td, "ref.fa"), "w") as fh:
                for n, c in zip(case["names"], case["codes"]):
                    fh.write(">%s\n%s\n" % (n, C2A[c].tobytes().decode()))
            p = test_flank.write_set_files(td, case["names"], case["j"], case["d"], case["ins"], case["f"])
            fa = subprocess.run([os.path.join(pyoracle.REF_DIR, "juncs_db"), str(case["min_anchor"]), str(case["max_seg_len"]), p["juncs"], p["ins"], p["dels"], p["fus"],
                                 os.path.join(td, "ref.fa")], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
        open(os.path.join(d, "juncs_db.fa"), "w").write(fa)
        # placements by exhaustion over the REFERENCE's contigs (parsed back from its FASTA)
        seqs = [l for l in fa.splitlines() if not l.startswith(">")]
        codes = [np.select([np.frombuffer(s.encode(), np.uint8) == x for x in (65, 67, 71, 84)], [0, 1, 2, 3], 4).astype(np.uint8) for s in seqs]
        hits = flank_oracle.search(codes, case["reads"], case["seg_bounds"], v_, max_hits, bool(npol_))
        np.savetxt(os.path.join(d, "placements.tsv"), hits, fmt="%d", delimiter="\t", header="read\tseg\tcontig\tpos\tantisense\tmismatches")
        json.dump({"case": cname, "seed": seed, "max_mismatches": v_, "max_multihits": max_hits, "ref_n_is_mismatch": npol_, "seg_bounds": [int(x) for x in bounds],
                   "n_reads": n_reads, "kwargs": kw, "contigs": len(seqs), "placements": int(len(hits)),
                   "made_by": "scripts/make_flank_golden.py with oracle/_ref/juncs_db (reference source, unmodified)"},
                  open(os.path.join(d, "config.json"), "w"), indent=1)
        print(gname, len(seqs), "contigs", len(hits), "placements")


if __name__ == "__main__":
    main()
