"""CPU tests of the C-ABI library: it loads, exports every symbol include/tophat_b200.h declares, its host
helpers agree with the Python packers, and it fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from tophat_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tophat_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(thb_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(built_library):
    lib = C.CDLL(built_library)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libtophat_b200.so does not export %s" % n


def test_params_default_matches_python(built_library):
    lib = capi.load_library()
    p = capi.Params()
    lib.thb_params_default(C.byref(p))
    q = capi.default_params()
    for name, _ in capi.Params._fields_[:-1]:
        assert getattr(p, name) == getattr(q, name), name


def test_pack_helpers_match_python(built_library):
    lib = capi.load_library()
    rng = np.random.default_rng(3)
    seq = bytes(np.frombuffer(b"ACGTNacgtnRYU", dtype=np.uint8)[rng.integers(0, 13, 300)])
    codes = synth.codes_from_ascii(seq)
    img = synth.build_ref_image(["c"], [codes])
    planes = np.zeros_like(img.planes); nmask = np.zeros_like(img.nmask)
    lib.thb_pack_bases(seq, len(seq), 0, planes.ctypes.data, nmask.ctypes.data)
    assert (planes == img.planes).all() and (nmask == img.nmask).all()
    read = bytes(np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, 101)])
    out = np.zeros(6, dtype="<u8")
    lib.thb_pack_read(read, 101, 2, out.ctypes.data)
    want = synth.pack_reads(synth.codes_from_ascii(read)[None, :], 2)[0]
    assert (out == want).all()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built_library):
    with pytest.raises(capi.ThbError):
        capi.Context(0)


def test_version(built_library):
    assert b"sm_100a" in capi.load_library().thb_version()


def test_join_pack_hits_matches_python(built_library):
    """thb_join_pack_hits (host helper: unpacked hits of one read -> 16-byte wire records + CIGAR side records) against the
    vectorised packer the bench uses; also its error for a read with more than 256 multi-op hits."""
    lib = capi.load_library()
    rng = np.random.default_rng(5)
    n = 400
    full = np.zeros(n, dtype=synth.JHIT_FULL_DTYPE)
    full["ref_id"] = rng.integers(1, 4, n); full["left"] = rng.integers(0, 1 << 28, n)
    full["flags"] = rng.integers(0, 8, n); full["mismatches"] = rng.integers(0, 4, n); full["splice_mms"] = rng.integers(0, 3, n)
    nops = np.where(rng.random(n) < 0.7, 1, rng.integers(1, 10, n)); full["n_ops"] = nops
    codes = np.array([1, 3, 5, 11, 13], dtype=np.uint32)
    for i in range(n):
        c = codes[rng.integers(0, 5, nops[i])]
        if nops[i] == 1 and rng.random() < 0.9:
            c[0] = 1
        full["ops"][i, :nops[i]] = (rng.integers(1, 3000, nops[i]).astype(np.uint32) << 4) | c
    # three "reads": hits [0,150), [150,151), [151,400)
    begins = np.array([0, 150, 151], dtype=np.int64)
    heads_py, ext_py, ops_begin = synth.pack_join_hits(full, begins)
    heads = np.zeros(n, dtype=synth.JHIT_DTYPE); ext = np.zeros(n, dtype=synth.JOPS_DTYPE); ne = 0
    ends = list(begins[1:]) + [n]
    for b, e in zip(begins, ends):
        assert ops_begin[list(begins).index(b)] == ne
        r = lib.thb_join_pack_hits(full[b:e].ctypes.data, int(e - b), heads[b:].ctypes.data, ext[ne:].ctypes.data)
        assert r >= 0
        ne += r
    assert ne == len(ext_py)
    assert (heads == heads_py).all() and (ext[:ne] == ext_py).all()
    assert ((heads["flags_nops"] & synth.JHIT_ONE_MATCH) != 0).sum() > 200 and ne > 50
    # more than 256 hits with a multi-op CIGAR in one read
    many = np.zeros(300, dtype=synth.JHIT_FULL_DTYPE); many["n_ops"] = 3
    many["ops"][:, 0] = (10 << 4) | 1; many["ops"][:, 1] = (100 << 4) | 11; many["ops"][:, 2] = (15 << 4) | 1
    h2 = np.zeros(300, dtype=synth.JHIT_DTYPE); e2 = np.zeros(300, dtype=synth.JOPS_DTYPE)
    assert lib.thb_join_pack_hits(many.ctypes.data, 300, h2.ctypes.data, e2.ctypes.data) == -5       # THB_EUNSUPPORTED
    with pytest.raises(ValueError):
        synth.pack_join_hits(many, np.array([0], dtype=np.int64))
