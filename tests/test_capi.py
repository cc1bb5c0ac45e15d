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
