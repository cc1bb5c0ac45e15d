"""ctypes mirror of include/tophat_b200.h and loader of libtophat_b200.so.

This is the Python face of the C ABI: the same structs, the same entry points, the same error
behaviour (negative return code -> ThbError carrying thb_last_error()).  There is no CPU fallback:
if the shared library was not built, or no sm_100 device is usable, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import synth

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtophat_b200.so")


class ThbError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "segment_length", "segment_mismatches", "min_segment_intron_length", "max_segment_intron_length",
        "max_insertion_length", "max_deletion_length", "max_seg_multihits", "inner_dist_mean",
        "inner_dist_std_dev", "bowtie2", "library_type", "fusion_search", "fusion_anchor_length",
        "fusion_min_dist", "max_report_intron_length", "min_report_intron_length", "min_anchor_len",
        "read_mismatches", "read_gap_length", "read_edit_dist", "bowtie2_max_penalty", "bowtie2_min_penalty",
        "bowtie2_penalty_for_N", "bowtie2_read_gap_open", "bowtie2_read_gap_cont", "bowtie2_ref_gap_open",
        "bowtie2_ref_gap_cont")] + [("reserved", C.c_int32 * 5)]


def default_params(**over) -> Params:
    """The reference binaries' defaults (common.cpp:79-180)."""
    p = Params()
    d = dict(segment_length=25, segment_mismatches=2, min_segment_intron_length=50,
             max_segment_intron_length=500000, max_insertion_length=3, max_deletion_length=3,
             max_seg_multihits=40, inner_dist_mean=200, inner_dist_std_dev=20, bowtie2=1, library_type=0,
             fusion_search=0, fusion_anchor_length=20, fusion_min_dist=10000000,
             max_report_intron_length=500000, min_report_intron_length=50, min_anchor_len=8,
             read_mismatches=2, read_gap_length=2, read_edit_dist=2, bowtie2_max_penalty=6,
             bowtie2_min_penalty=2, bowtie2_penalty_for_N=1, bowtie2_read_gap_open=5, bowtie2_read_gap_cont=3,
             bowtie2_ref_gap_open=5, bowtie2_ref_gap_cont=3)
    d.update(over)
    for k, v in d.items():
        setattr(p, k, int(v))
    return p


class RefImageC(C.Structure):
    _fields_ = [("n_contigs", C.c_uint32), ("contig_start", C.c_void_p), ("contig_len", C.c_void_p),
                ("n_blocks", C.c_uint64), ("planes", C.c_void_p), ("nmask", C.c_void_p)]


class BatchC(C.Structure):
    _fields_ = [("n_bundles", C.c_uint32), ("n_segs", C.c_uint32), ("read_words", C.c_uint32),
                ("reserved", C.c_uint32), ("bundles", C.c_void_p), ("seg_count", C.c_void_p),
                ("reads", C.c_void_p), ("n_hits", C.c_uint64), ("hits", C.c_void_p),
                ("n_partner_hits", C.c_uint64), ("partner_hits", C.c_void_p), ("order_base", C.c_uint64)]


class JoinBatchC(C.Structure):
    _fields_ = [("n_bundles", C.c_uint32), ("n_segs", C.c_uint32), ("read_words", C.c_uint32), ("reserved", C.c_uint32),
                ("bundles", C.c_void_p), ("seg_count", C.c_void_p), ("reads", C.c_void_p), ("n_hits", C.c_uint64), ("hits", C.c_void_p),
                ("n_ops_ext", C.c_uint64), ("ops_ext", C.c_void_p)]


class JoinTimingC(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float), ("enum_ms", C.c_float), ("merge_ms", C.c_float), ("launches", C.c_uint32),
                ("n_chains", C.c_uint64), ("n_closures", C.c_uint64), ("n_joined", C.c_uint64), ("algorithmic_bytes", C.c_uint64),
                ("n_simple_chains", C.c_uint64), ("n_abutting_chains", C.c_uint64),
                ("merge_simple_ms", C.c_float), ("merge_abutting_ms", C.c_float), ("merge_general_ms", C.c_float), ("begin_ms", C.c_float)]


def join_batch_c(b: "synth.PackedJoinBatch") -> JoinBatchC:
    s = JoinBatchC()
    s.n_bundles = b.n_bundles; s.n_segs = b.n_segs; s.read_words = b.read_words
    s.bundles = b.bundles.ctypes.data; s.seg_count = b.seg_count.ctypes.data; s.reads = b.reads.ctypes.data
    s.n_hits = b.hits.shape[0]; s.hits = b.hits.ctypes.data
    s.n_ops_ext = b.ops_ext.shape[0]; s.ops_ext = b.ops_ext.ctypes.data
    return s


class ResultsC(C.Structure):
    _fields_ = [("n_junctions", C.c_uint64), ("junctions", C.c_void_p),
                ("n_deletions", C.c_uint64), ("deletions", C.c_void_p),
                ("n_insertions", C.c_uint64), ("insertions", C.c_void_p),
                ("n_fusions", C.c_uint64), ("fusions", C.c_void_p)]


class TimingC(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("scan_kernel_ms", C.c_float), ("finish_ms", C.c_float),
                ("total_ms", C.c_float), ("bundle_ms", C.c_float), ("hit_ms", C.c_float), ("rescue_ms", C.c_float),
                ("rescued_windows_ms", C.c_float), ("window_scan_ms", C.c_float), ("indel_ms", C.c_float),
                ("fusion_enum_ms", C.c_float), ("fusion_detect_ms", C.c_float), ("n_windows", C.c_uint64), ("n_indel_tasks", C.c_uint64),
                ("n_rescue_tasks", C.c_uint64), ("n_juncs_emitted", C.c_uint64), ("n_fusion_tasks", C.c_uint64),
                ("algorithmic_bytes", C.c_uint64), ("kernel_launches", C.c_uint32), ("total_launches", C.c_uint32)]


def ref_image_c(ref: synth.RefImage) -> RefImageC:
    """Borrowed-pointer view; `ref` must outlive the struct."""
    r = RefImageC()
    r.n_contigs = len(ref.names)
    r.contig_start = ref.contig_start.ctypes.data
    r.contig_len = ref.contig_len.ctypes.data
    r.n_blocks = ref.n_blocks
    r.planes = ref.planes.ctypes.data
    r.nmask = ref.nmask.ctypes.data
    return r


def batch_c(b: synth.PackedBatch) -> BatchC:
    """Borrowed-pointer view of host numpy arrays; `b` must outlive the struct."""
    s = BatchC()
    s.n_bundles = b.n_bundles
    s.n_segs = b.n_segs
    s.read_words = b.read_words
    s.bundles = b.bundles.ctypes.data
    s.seg_count = b.seg_count.ctypes.data
    s.reads = b.reads.ctypes.data
    s.n_hits = b.hits.shape[0]
    s.hits = b.hits.ctypes.data
    s.n_partner_hits = b.partner_hits.shape[0]
    s.partner_hits = b.partner_hits.ctypes.data
    s.order_base = b.order_base
    return s


def _copy_records(ptr: Optional[int], n: int, dtype: np.dtype, copy: bool = True) -> np.ndarray:
    if not n:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    a = np.frombuffer(buf, dtype=dtype, count=n)
    return a.copy() if copy else a


class SegJuncsResults:
    def __init__(self, juncs, dels, ins, fus):
        self.junctions, self.deletions, self.insertions, self.fusions = juncs, dels, ins, fus

    @classmethod
    def from_c(cls, r: ResultsC, copy: bool = True) -> "SegJuncsResults":
        """copy=False: views of the context's own (page-locked) result arrays, valid until the next thb_segjuncs_begin."""
        return cls(_copy_records(r.junctions, r.n_junctions, synth.JUNCTION_DTYPE, copy),
                   _copy_records(r.deletions, r.n_deletions, synth.JUNCTION_DTYPE, copy),
                   _copy_records(r.insertions, r.n_insertions, synth.INSERTION_DTYPE, copy),
                   _copy_records(r.fusions, r.n_fusions, synth.FUSION_DTYPE, copy))


_lib = None


THB_MAX_SEGS = 12
FLANK_CONTIG_DTYPE = np.dtype([("kind", "<u4"), ("ref_id", "<u4"), ("ref_id2", "<u4"), ("left_start", "<u4"), ("left", "<u4"), ("right", "<u4"),
                               ("right_end", "<u4"), ("aux", "<u4"), ("length", "<u4"), ("ins_seq", "S20")])
FLANK_HIT_DTYPE = np.dtype([("read", "<u4"), ("contig", "<u4"), ("seg", "u1"), ("pos", "u1"), ("antisense", "u1"), ("mismatches", "u1")])
assert FLANK_CONTIG_DTYPE.itemsize == 56 and FLANK_HIT_DTYPE.itemsize == 12


class FlankParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("max_mismatches", "max_multihits", "min_seg_len", "max_seg_len", "min_anchor", "ref_n_is_mismatch")]


class FlankBatchC(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("read_words", C.c_uint32), ("n_segs", C.c_uint32), ("reserved", C.c_uint32),
                ("reads", C.c_void_p), ("seg_bounds", C.c_uint16 * (THB_MAX_SEGS + 1))]


class FlankTimingC(C.Structure):
    _fields_ = [("index_ms", C.c_float), ("h2d_ms", C.c_float), ("match_ms", C.c_float), ("post_ms", C.c_float), ("d2h_ms", C.c_float),
                ("n_contigs", C.c_uint64), ("n_index_entries", C.c_uint64), ("n_verified", C.c_uint64), ("n_hits", C.c_uint64),
                ("algorithmic_bytes", C.c_uint64), ("launches", C.c_uint32), ("reserved", C.c_uint32)]


def flank_batch_c(reads_ptr: int, n_reads: int, read_words: int, seg_bounds) -> FlankBatchC:
    b = FlankBatchC()
    b.n_reads = n_reads; b.read_words = read_words; b.n_segs = len(seg_bounds) - 1; b.reads = reads_ptr
    for i, v in enumerate(seg_bounds):
        b.seg_bounds[i] = int(v)
    return b


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Loads libtophat_b200.so; raises ThbError (never falls back) if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ThbError("libtophat_b200.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`" % p)
    lib = C.CDLL(p)
    lib.thb_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.thb_destroy.argtypes = [C.c_void_p]
    lib.thb_destroy.restype = None
    lib.thb_last_error.argtypes = [C.c_void_p]
    lib.thb_last_error.restype = C.c_char_p
    lib.thb_version.restype = C.c_char_p
    lib.thb_params_default.argtypes = [C.POINTER(Params)]
    lib.thb_params_default.restype = None
    lib.thb_ref_upload.argtypes = [C.c_void_p, C.POINTER(RefImageC)]
    lib.thb_segjuncs_begin.argtypes = [C.c_void_p, C.POINTER(Params)]
    lib.thb_segjuncs_fusion_ignore.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.thb_segjuncs_submit.argtypes = [C.c_void_p, C.POINTER(BatchC)]
    lib.thb_segjuncs_submit_device.argtypes = [C.c_void_p, C.POINTER(BatchC)]
    lib.thb_segjuncs_finish.argtypes = [C.c_void_p, C.POINTER(ResultsC)]
    lib.thb_last_timing.argtypes = [C.c_void_p, C.POINTER(TimingC)]
    lib.thb_segjuncs_finish_resident.argtypes = [C.c_void_p, C.POINTER(ResultsC)]
    lib.thb_segjuncs_fetch.argtypes = [C.c_void_p]
    lib.thb_join_begin_resident.argtypes = [C.c_void_p, C.POINTER(Params)]
    lib.thb_stream.argtypes = [C.c_void_p]
    lib.thb_stream.restype = C.c_void_p
    lib.thb_pack_bases.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.thb_pack_bases.restype = None
    lib.thb_pack_read.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.thb_pack_read.restype = None
    lib.thb_join_pack_hits.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.thb_join_set_fusions.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.thb_join_begin.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.thb_join_submit.argtypes = [C.c_void_p, C.POINTER(JoinBatchC), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.thb_join_submit_device.argtypes = [C.c_void_p, C.POINTER(JoinBatchC), C.POINTER(C.c_uint64)]
    lib.thb_join_fetch.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.thb_join_last_timing.argtypes = [C.c_void_p, C.POINTER(JoinTimingC)]
    lib.thb_flank_begin.argtypes = [C.c_void_p, C.POINTER(FlankParams), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.thb_flank_contigs.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.thb_flank_submit.argtypes = [C.c_void_p, C.POINTER(FlankBatchC), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.thb_flank_submit_device.argtypes = [C.c_void_p, C.POINTER(FlankBatchC), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.thb_flank_last_timing.argtypes = [C.c_void_p, C.POINTER(FlankTimingC)]
    lib.thb_flank_spliced_hits.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.thb_alloc_pinned.argtypes = [C.c_size_t]
    lib.thb_alloc_pinned.restype = C.c_void_p
    lib.thb_free_pinned.argtypes = [C.c_void_p]
    lib.thb_free_pinned.restype = None
    lib.thb_nccl_unique_id.argtypes = [C.c_void_p]
    lib.thb_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.thb_segjuncs_allgather.argtypes = [C.c_void_p]
    if path is None:
        _lib = lib
    return lib


class Context:
    """One device context (thb_ctx).  Mirrors the call sequence of the host binaries."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.thb_create(device, C.byref(h))
        if rc != 0:
            msg = self.lib.thb_last_error(None)
            raise ThbError("thb_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.h = h
        self._keep = []

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.lib.thb_last_error(self.h)
            raise ThbError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))

    def close(self) -> None:
        if self.h:
            self.lib.thb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ref_upload(self, ref: synth.RefImage) -> None:
        img = ref_image_c(ref)
        self._check(self.lib.thb_ref_upload(self.h, C.byref(img)), "thb_ref_upload")

    def segjuncs_begin(self, params: Params) -> None:
        self._check(self.lib.thb_segjuncs_begin(self.h, C.byref(params)), "thb_segjuncs_begin")

    def segjuncs_fusion_ignore(self, ref_ids) -> None:
        a = np.ascontiguousarray(np.asarray(ref_ids, dtype="<u4"))
        self._check(self.lib.thb_segjuncs_fusion_ignore(self.h, a.ctypes.data if a.size else None, a.size), "thb_segjuncs_fusion_ignore")

    def segjuncs_submit(self, batch: synth.PackedBatch) -> None:
        b = batch_c(batch)
        self._check(self.lib.thb_segjuncs_submit(self.h, C.byref(b)), "thb_segjuncs_submit")

    def segjuncs_submit_device(self, b: BatchC) -> None:
        self._check(self.lib.thb_segjuncs_submit_device(self.h, C.byref(b)), "thb_segjuncs_submit_device")

    def segjuncs_finish(self, copy: bool = True) -> SegJuncsResults:
        r = ResultsC()
        self._check(self.lib.thb_segjuncs_finish(self.h, C.byref(r)), "thb_segjuncs_finish")
        return SegJuncsResults.from_c(r, copy)

    def segjuncs_finish_raw(self) -> ResultsC:
        """The C struct itself (counts + pointers into the context's page-locked arrays): no numpy objects are built."""
        r = ResultsC()
        self._check(self.lib.thb_segjuncs_finish(self.h, C.byref(r)), "thb_segjuncs_finish")
        return r

    def segjuncs_finish_resident(self) -> ResultsC:
        """Sets built and kept on the device, counts filled, arrays downloading in the background: valid after segjuncs_fetch()."""
        r = ResultsC()
        self._check(self.lib.thb_segjuncs_finish_resident(self.h, C.byref(r)), "thb_segjuncs_finish_resident")
        return r

    def segjuncs_fetch(self) -> None:
        self._check(self.lib.thb_segjuncs_fetch(self.h), "thb_segjuncs_fetch")

    def join_begin_resident(self, params: Params) -> None:
        """thb_join_begin with the device-resident sets of the segment_juncs pass this context has just finished."""
        self._check(self.lib.thb_join_begin_resident(self.h, C.byref(params)), "thb_join_begin_resident")

    def timing(self) -> TimingC:
        t = TimingC()
        self._check(self.lib.thb_last_timing(self.h, C.byref(t)), "thb_last_timing")
        return t

    def stream(self) -> int:
        return int(self.lib.thb_stream(self.h) or 0)

    # ---- multi-GPU exchange (NCCL)
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.thb_nccl_unique_id(buf)
        if rc != 0:
            raise ThbError("thb_nccl_unique_id failed (%d)" % rc)
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, world: int) -> None:
        buf = C.create_string_buffer(uid, 128)
        self._check(self.lib.thb_comm_init(self.h, buf, rank, world), "thb_comm_init")

    def segjuncs_allgather(self) -> None:
        self._check(self.lib.thb_segjuncs_allgather(self.h), "thb_segjuncs_allgather")

    # ---- long_spanning_reads join
    def join_begin(self, params: Params, junctions: np.ndarray, insertions: np.ndarray) -> None:
        """junctions: JUNCTION_DTYPE sorted unique (incl. deletions as (ref, left-1, right)); insertions: INSERTION_DTYPE sorted."""
        self._jkeep = (np.ascontiguousarray(junctions), np.ascontiguousarray(insertions))
        j, i = self._jkeep
        self._check(self.lib.thb_join_begin(self.h, C.byref(params), j.ctypes.data if j.size else None, j.shape[0],
                                            i.ctypes.data if i.size else None, i.shape[0]), "thb_join_begin")

    def join_set_fusions(self, fusions: np.ndarray) -> None:
        """--fusion-search: FUSION_DTYPE records sorted unique in Fusion order (fusions.h:40-70); call after join_begin."""
        self._fkeep = np.ascontiguousarray(fusions)
        f = self._fkeep
        self._check(self.lib.thb_join_set_fusions(self.h, f.ctypes.data if f.size else None, f.shape[0]), "thb_join_set_fusions")

    def join_submit(self, batch) -> np.ndarray:
        b = batch if isinstance(batch, JoinBatchC) else join_batch_c(batch)
        out = C.c_void_p(); n = C.c_uint64()
        self._check(self.lib.thb_join_submit(self.h, C.byref(b), C.byref(out), C.byref(n)), "thb_join_submit")
        return _copy_records(out.value, n.value, synth.JOINED_DTYPE)

    def join_submit_device(self, b: JoinBatchC) -> int:
        n = C.c_uint64()
        self._check(self.lib.thb_join_submit_device(self.h, C.byref(b), C.byref(n)), "thb_join_submit_device")
        return int(n.value)

    def join_fetch(self) -> np.ndarray:
        out = C.c_void_p(); n = C.c_uint64()
        self._check(self.lib.thb_join_fetch(self.h, C.byref(out), C.byref(n)), "thb_join_fetch")
        return _copy_records(out.value, n.value, synth.JOINED_DTYPE)

    def join_timing(self) -> JoinTimingC:
        t = JoinTimingC()
        self._check(self.lib.thb_join_last_timing(self.h, C.byref(t)), "thb_join_last_timing")
        return t

    # ---- junction-flank matcher (juncs_db + bowtie-build + bowtie of the segments, tophat.py:2546-2600, 3686-3741) ----
    def flank_begin(self, params: FlankParams, junctions: np.ndarray, deletions: np.ndarray, insertions: np.ndarray, fusions: np.ndarray) -> None:
        keep = [np.ascontiguousarray(a) for a in (junctions, deletions, insertions, fusions)]
        args = []
        for a in keep:
            args += [a.ctypes.data if a.size else None, a.shape[0]]
        self._check(self.lib.thb_flank_begin(self.h, C.byref(params), *args), "thb_flank_begin")

    def flank_contigs(self) -> np.ndarray:
        out = C.c_void_p(); n = C.c_uint64()
        self._check(self.lib.thb_flank_contigs(self.h, C.byref(out), C.byref(n)), "thb_flank_contigs")
        return _copy_records(out.value, n.value, FLANK_CONTIG_DTYPE)

    def flank_submit(self, reads: np.ndarray, read_words: int, seg_bounds, device_ptr: Optional[int] = None, copy: bool = True) -> np.ndarray:
        """reads: (n, 3*read_words) uint64 bit planes (synth.pack_reads); or device_ptr + reads = number of reads."""
        out = C.c_void_p(); n = C.c_uint64()
        if device_ptr is None:
            r = np.ascontiguousarray(reads)
            b = flank_batch_c(r.ctypes.data if r.size else None, r.shape[0], read_words, seg_bounds)
            self._check(self.lib.thb_flank_submit(self.h, C.byref(b), C.byref(out), C.byref(n)), "thb_flank_submit")
        else:
            b = flank_batch_c(device_ptr, int(reads), read_words, seg_bounds)
            self._check(self.lib.thb_flank_submit_device(self.h, C.byref(b), C.byref(out), C.byref(n)), "thb_flank_submit_device")
        return _copy_records(out.value, n.value, FLANK_HIT_DTYPE, copy)

    def flank_spliced_hits(self, min_anchor_len: int, copy: bool = True) -> np.ndarray:
        """JHIT_FULL_DTYPE records, one per placement of the last flank_submit (n_ops == 0: discarded by the reference's rules)."""
        out = C.c_void_p(); n = C.c_uint64()
        self._check(self.lib.thb_flank_spliced_hits(self.h, min_anchor_len, C.byref(out), C.byref(n)), "thb_flank_spliced_hits")
        return _copy_records(out.value, n.value, synth.JHIT_FULL_DTYPE, copy)

    def flank_timing(self) -> FlankTimingC:
        t = FlankTimingC()
        self._check(self.lib.thb_flank_last_timing(self.h, C.byref(t)), "thb_flank_last_timing")
        return t


def join_sets_from_results(res: "SegJuncsResults"):
    """The std::set<Junction> long_spanning_reads builds from segment.juncs + segment.deletions (deletions enter as
    Junction(ref, left - 1 + 1 - 1 ... ) i.e. exactly the Deletion record: long_spanning_reads.cpp:2916-2944) and the insertion set."""
    j = np.concatenate([res.junctions, res.deletions]) if res.deletions.size else res.junctions.copy()
    if res.deletions.size:
        j[len(res.junctions):]["antisense"] = 0
    order = np.lexsort((j["antisense"], j["right"], j["left"], j["ref_id"]))
    j = j[order]
    if j.size:
        keep = np.ones(j.size, bool)
        keep[1:] = ~((j["ref_id"][1:] == j["ref_id"][:-1]) & (j["left"][1:] == j["left"][:-1]) & (j["right"][1:] == j["right"][:-1]) &
                     (j["antisense"][1:] == j["antisense"][:-1]))
        j = j[keep]
    return np.ascontiguousarray(j), np.ascontiguousarray(res.insertions)
