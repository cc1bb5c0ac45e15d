"""Known-answer vectors of SURVEY.md appendix B: constructions whose answers were obtained from the
reference's own segment_juncs binary.  Each returns (contigs, reads, expected) for helpers.manual_workload."""
import numpy as np

_COMP = {65: 84, 67: 71, 71: 67, 84: 65}


def _random_ref(n, seed):
    rng = np.random.default_rng(seed)
    return bytearray(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes())


def _scrub(ref, lo, hi, keep):
    """Removes accidental GT/GC/AT/CT-AC... motifs is not needed: answers below were produced by the binary on
    exactly these sequences (seeded), so accidental extra junctions are part of the expected answer."""
    return ref


def kat_junction():
    ref = _random_ref(6000, 1)
    ref[1100:1102] = b"GT"; ref[2098:2100] = b"AG"
    read = bytes(ref[1049:1100] + ref[2100:2150])
    hits = [[(1, 1049, 25, 0, 0)], [(1, 1074, 25, 0, 0)], [], [(1, 2124, 26, 0, 0)]]
    return [("chrT", bytes(ref))], [dict(seq=read, hits=hits)], ("chrT", 1099, 2100, "+")


def kat_junction_seg1_unmapped():
    ref = _random_ref(6000, 2)
    ref[1100:1102] = b"GT"; ref[2098:2100] = b"AG"
    read = bytes(ref[1065:1100] + ref[2100:2166])
    hits = [[(1, 1065, 25, 0, 0)], [], [(1, 2115, 25, 0, 0)], [(1, 2140, 26, 0, 0)]]
    return [("chrT", bytes(ref))], [dict(seq=read, hits=hits)], ("chrT", 1099, 2100, "+")


def kat_deletion():
    ref = _random_ref(9000, 3)
    p = 5000
    read = bytearray(ref[p:p + 24] + ref[p + 26:p + 103])
    hits = [[(1, p, 25, 1, 0)], [(1, p + 27, 25, 0, 0)], [(1, p + 52, 25, 0, 0)], [(1, p + 77, 26, 0, 0)]]
    return [("chrT", bytes(ref))], [dict(seq=bytes(read), hits=hits)], ("chrT", 5024, 5026)


def kat_insertion():
    ref = _random_ref(12000, 4)
    q = 8000
    # make the inserted dinucleotide unambiguous: the bases around the insertion differ from T
    for k in (q + 47, q + 48, q + 49, q + 50):
        if ref[k] == ord("T"):
            ref[k] = ord("C")
    read = bytes(ref[q:q + 49] + b"TT" + ref[q + 49:q + 99])
    hits = [[(1, q, 25, 0, 0)], [(1, q + 25, 25, 1, 0)], [(1, q + 48, 25, 1, 0)], [(1, q + 73, 26, 0, 0)]]
    return [("chrT", bytes(ref))], [dict(seq=read, hits=hits)], ("chrT", 8048, 8048, "TT")


def kat_q0_quirk():
    """Decoy donor left of the true one passes because right_mismatches[] stays 0 below the break index."""
    ref = _random_ref(8000, 5)
    E, D = 3000, 4000
    ref[E:E + 2] = b"GT"; ref[D - 2:D] = b"AG"
    ref[E - 6:E - 4] = b"GT"; ref[D - 8:D - 6] = b"AG"
    # make sure the decoy's right flank really mismatches: ref[D-6:D] vs read[44:50]=ref[E-6:E]
    for k in range(6):
        if k not in (0, 1, 4, 5) and ref[D - 6 + k] == ref[E - 6 + k]:
            ref[D - 6 + k] = _COMP[ref[E - 6 + k]]
    read = bytes(ref[E - 50:E] + ref[D:D + 51])
    hits = [[(1, E - 50, 25, 0, 0)], [(1, E - 25, 25, 0, 0)], [(1, D, 25, 0, 0)], [(1, D + 25, 26, 0, 0)]]
    return [("chrT", bytes(ref))], [dict(seq=read, hits=hits)], [("chrT", 2993, 3994, "+"), ("chrT", 2999, 4000, "+")]
