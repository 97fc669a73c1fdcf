"""ctypes binding of libnanocaller_b200.so (include/nanocaller_b200.h).

There is no CPU fallback: if the CUDA library cannot be built/loaded, or no sm_100 device is
present, constructing a `Context` raises."""
import ctypes
import os

import numpy as np

from .. import build as _build

NC_OK, NC_ECUDA, NC_EINVAL, NC_ESTATE, NC_ENOMEM, NC_EOVERFLOW = 0, -1, -2, -3, -4, -5
SEQ_CODES = {"ont": 0, "short_ont": 1, "ul_ont": 2, "ul_ont_extreme": 3, "pacbio": 4}
SITE_ELEMS, SITE_STRIDE = 1025, 1032

EXPORTS = ["nc_abi_version", "nc_create", "nc_destroy", "nc_last_error", "nc_sync", "nc_set_blocking_sync", "nc_get_timings",
           "nc_device_sm_count", "nc_event_record", "nc_event_elapsed_ms", "nc_invalidate_decode", "nc_stage_reads", "nc_decode_reads", "nc_snp_scan", "nc_snp_fetch", "nc_snp_fetch_range", "nc_indel_fetch_range",
           "nc_load_snp_weights", "nc_snp_forward", "nc_snp_fetch_probs", "nc_snp_model_forward", "nc_snp_device_buffers",
           "nc_load_indel_weights", "nc_indel_model_forward", "nc_stage_tags", "nc_indel_scan", "nc_indel_fetch_variants",
           "nc_indel_build", "nc_indel_fetch", "nc_indel_forward", "nc_indel_fetch_probs", "nc_indel_fetch_alleles",
           "nc_bam_device_open", "nc_bam_device_contig", "nc_bam_device_stage", "nc_bam_device_close", "nc_bam_device_timings", "nc_bam_device_walk_mode", "nc_get_indel_timings", "nc_nw_trace", "nc_allele_predict_batch",
           "nc_format_snp_records"]


class NcSnpParams(ctypes.Structure):
    _fields_ = [("thr_lo", ctypes.c_double), ("thr_hi", ctypes.c_double), ("min_allele_freq", ctypes.c_double),
                ("mincov", ctypes.c_int32), ("maxcov", ctypes.c_int32), ("min_nbr_sites", ctypes.c_int32),
                ("seq", ctypes.c_int32), ("supplementary", ctypes.c_int32), ("haploid", ctypes.c_int32)]


class NcTimings(ctypes.Structure):
    _fields_ = [("decode_ms", ctypes.c_float), ("scan_ms", ctypes.c_float), ("tensor_ms", ctypes.c_float),
                ("cnn_ms", ctypes.c_float), ("launches", ctypes.c_uint64), ("tensor_bytes", ctypes.c_uint64),
                ("scan_bytes", ctypes.c_uint64), ("cnn_a_ms", ctypes.c_float), ("reserved", ctypes.c_float)]


class NcIndelParams(ctypes.Structure):
    _fields_ = [("ins_t", ctypes.c_double), ("del_t", ctypes.c_double), ("mincov", ctypes.c_int32), ("maxcov", ctypes.c_int32),
                ("win_size", ctypes.c_int32), ("small_win_size", ctypes.c_int32), ("window_after", ctypes.c_int32),
                ("supplementary", ctypes.c_int32), ("haploid", ctypes.c_int32), ("impute_indel_phase", ctypes.c_int32)]


class NcIndelTimings(ctypes.Structure):
    _fields_ = [("scan_ms", ctypes.c_float), ("reads_ms", ctypes.c_float), ("align_ms", ctypes.c_float), ("msa_ms", ctypes.c_float),
                ("cnn_ms", ctypes.c_float), ("allele_ms", ctypes.c_float), ("reserved", ctypes.c_float * 2), ("n_sites", ctypes.c_uint64), ("n_entries", ctypes.c_uint64),
                ("scan_bytes", ctypes.c_uint64), ("build_bytes", ctypes.c_uint64)]


class NcBamDeviceContig(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 256), ("length", ctypes.c_int32), ("reserved", ctypes.c_int32), ("n_reads", ctypes.c_int64),
                ("n_tagged", ctypes.c_int64)]


VARIANT_DTYPE = np.dtype([("key", "<i4"), ("type", "<i4"), ("chunk", "<i4"), ("src", "<i4")])
INDEL_META_DTYPE = np.dtype([("pos", "<i4"), ("chunk", "<i4"), ("type", "<i4"), ("phase", "<i4"), ("ref_len", "<i4"),
                             ("n", "<i4", (3,)), ("cns_len", "<i4", (3,)), ("ok", "<i4", (3,))])
INDEL_CNS_MAX = 544
CHUNK_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4")])
META_DTYPE = np.dtype([("pos", "<i4"), ("chunk", "<i4"), ("dp", "<i4"), ("alt", "<i4"), ("fwd", "<u2", (4,)),
                       ("rev", "<u2", (4,)), ("ref_code", "u1"), ("n_left", "u1"), ("n_right", "u1"),
                       ("reserved", "u1"), ("sample_depth", "<i4")])
assert META_DTYPE.itemsize == 40

_lib = None


class NcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libnanocaller_b200 error %d: %s" % (code, msg))
        self.code = code


def library_path():
    return _build.LIB_CUDA


def load_library():
    """Build (if stale and nvcc is present) and load the CUDA library; raises when impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build_cuda()
    lib = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.nc_abi_version.restype = ctypes.c_int
    lib.nc_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    lib.nc_destroy.argtypes = [vp]
    lib.nc_destroy.restype = None
    lib.nc_last_error.argtypes = [vp]
    lib.nc_last_error.restype = ctypes.c_char_p
    lib.nc_sync.argtypes = [vp]
    lib.nc_set_blocking_sync.argtypes = [vp, ctypes.c_int]
    lib.nc_get_timings.argtypes = [vp, ctypes.POINTER(NcTimings)]
    lib.nc_device_sm_count.argtypes = [vp]
    lib.nc_event_record.argtypes = [vp, ctypes.c_int]
    lib.nc_event_elapsed_ms.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    lib.nc_invalidate_decode.argtypes = [vp]
    lib.nc_stage_reads.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64]
    lib.nc_decode_reads.argtypes = [vp]
    lib.nc_snp_scan.argtypes = [vp, ctypes.POINTER(NcSnpParams), vp, i32, vp, i32, ctypes.POINTER(i64)]
    lib.nc_snp_fetch.argtypes = [vp, vp, vp, vp, vp]
    lib.nc_snp_fetch_range.argtypes = [vp, i64, i64, vp]
    lib.nc_indel_fetch_range.argtypes = [vp, i64, i64, vp]
    lib.nc_load_snp_weights.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_double, ctypes.c_int]
    lib.nc_snp_forward.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    lib.nc_snp_fetch_probs.argtypes = [vp, vp]
    lib.nc_snp_model_forward.argtypes = [vp, vp, vp, i64, ctypes.c_int, ctypes.c_int, vp]
    lib.nc_snp_device_buffers.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(i64)]
    lib.nc_load_indel_weights.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_int]
    lib.nc_indel_model_forward.argtypes = [vp, vp, i64, ctypes.c_int, ctypes.c_int, vp]
    lib.nc_stage_tags.argtypes = [vp, vp, vp]
    lib.nc_indel_scan.argtypes = [vp, ctypes.POINTER(NcIndelParams), vp, i32, vp, i32, ctypes.POINTER(i64)]
    lib.nc_indel_fetch_variants.argtypes = [vp, vp]
    lib.nc_indel_build.argtypes = [vp, ctypes.POINTER(NcIndelParams), vp, i32, vp, i64]
    lib.nc_indel_fetch.argtypes = [vp, vp, vp, vp]
    lib.nc_indel_forward.argtypes = [vp, ctypes.c_int, vp]
    lib.nc_indel_fetch_probs.argtypes = [vp, vp]
    lib.nc_indel_fetch_alleles.argtypes = [vp, vp]
    lib.nc_get_indel_timings.argtypes = [vp, ctypes.POINTER(NcIndelTimings)]
    lib.nc_bam_device_open.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(i32)]
    lib.nc_bam_device_contig.argtypes = [vp, i32, ctypes.POINTER(NcBamDeviceContig)]
    lib.nc_bam_device_stage.argtypes = [vp, i32, vp, i64, i64]
    lib.nc_bam_device_close.argtypes = [vp]
    lib.nc_bam_device_walk_mode.argtypes = [vp]
    lib.nc_bam_device_timings.argtypes = [vp, ctypes.POINTER(ctypes.c_float * 4), ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.nc_nw_trace.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, vp, i32]
    for name in EXPORTS:
        if name not in ("nc_destroy", "nc_last_error"):
            getattr(lib, name).restype = ctypes.c_int64 if name == "nc_format_snp_records" else ctypes.c_int
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def snp_params(dct, ploidy):
    """NcSnpParams from the reference's `dct` (generate_SNP_pileups.py:113-132,170-183,202,215,244)."""
    return NcSnpParams(float(dct["threshold"][0]), float(dct["threshold"][1]), float(dct["min_allele_freq"]),
                       int(dct["mincov"]), int(dct["maxcov"]), int(dct["min_nbr_sites"]), SEQ_CODES[dct["seq"]],
                       1 if dct.get("supplementary") else 0, 1 if ploidy == "haploid" else 0)


def indel_params(dct, haploid=False):
    """NcIndelParams from the reference's `dct` (generate_indel_pileups.py:136-157)."""
    return NcIndelParams(float(dct["ins_t"]), float(dct["del_t"]), int(dct["mincov"]), int(dct["maxcov"]), int(dct["win_size"]),
                         int(dct["small_win_size"]), 260 if dct["seq"] == "pacbio" else 160, 1 if dct.get("supplementary") else 0,
                         1 if haploid else 0, 1 if (dct.get("impute_indel_phase") and not haploid) else 0)


def nw_trace(query_codes, ref_codes, gap_open=9, gap_extend=1, match=20, mismatch=-10):
    """Host-side affine alignment of the library (replaces parasail.nw_trace): -> list of (op code, length), ops '='7 'X'8 'I'1 'D'2."""
    lib = load_library()
    q = np.ascontiguousarray(query_codes, np.uint8)
    r = np.ascontiguousarray(ref_codes, np.uint8)
    out = np.empty(len(q) + len(r) + 2, np.uint32)
    n = lib.nc_nw_trace(_p(q) if len(q) else None, len(q), _p(r) if len(r) else None, len(r), gap_open, gap_extend, match, mismatch, _p(out), len(out))
    if n < 0:
        raise NcError(n, "nc_nw_trace failed")
    return [(int(w & 15), int(w >> 4)) for w in out[:n]]


def allele_predict_batch(alt_flat, alt_off, alt_len, ref_flat, ref_off, ref_len, max_range, threads=0,
                         gap_open=9, gap_extend=1, match=20, mismatch=-10):
    """Batch of allele_prediction calls (generate_indel_pileups.py:77-127) on host threads -> (ref_out_len, alt_out_len) int32
    arrays, -1 where the reference returns (None, None)."""
    lib = load_library()
    n = len(alt_len)
    alt_flat = np.ascontiguousarray(alt_flat, np.uint8); ref_flat = np.ascontiguousarray(ref_flat, np.uint8)
    alt_off = np.ascontiguousarray(alt_off, np.int64); ref_off = np.ascontiguousarray(ref_off, np.int64)
    alt_len = np.ascontiguousarray(alt_len, np.int32); ref_len = np.ascontiguousarray(ref_len, np.int32)
    max_range = np.ascontiguousarray(max_range, np.int32)
    ro, ao = np.empty(n, np.int32), np.empty(n, np.int32)
    if n == 0:
        return ro, ao
    rc = lib.nc_allele_predict_batch(n, _p(alt_flat), _p(alt_off), _p(alt_len), _p(ref_flat), _p(ref_off), _p(ref_len), _p(max_range),
                                     gap_open, gap_extend, match, mismatch, int(threads), _p(ro), _p(ao))
    if rc != NC_OK:
        raise NcError(rc, "nc_allele_predict_batch failed")
    return ro, ao


def format_snp_records(chrom, pos, ref_code, probs, dp, alt, fwd, rev, haploid=False, threads=0):
    """Decision + record text of snpCaller.py:113-163 / :183-198 in the library (threaded).  -> (bytes blob, line_off int64[n+1],
    is_pass bool[n]); the records are blob[line_off[i]:line_off[i+1]]."""
    lib = load_library()
    n = len(pos)
    pos = np.ascontiguousarray(pos, np.int32); ref_code = np.ascontiguousarray(ref_code, np.uint8)
    probs = np.ascontiguousarray(probs, np.float32).reshape(n, 4)
    dp = np.ascontiguousarray(dp, np.int32); alt = np.ascontiguousarray(alt, np.int32)
    fwd = np.ascontiguousarray(fwd, np.uint16).reshape(n, 4); rev = np.ascontiguousarray(rev, np.uint16).reshape(n, 4)
    out = np.empty(max(1, n) * 512, np.uint8)
    off = np.zeros(n + 1, np.int64)
    ok = np.zeros(max(1, n), np.uint8)
    f = lib.nc_format_snp_records
    f.argtypes = [ctypes.c_char_p, ctypes.c_int64] + [ctypes.c_void_p] * 7 + [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    tot = f(chrom.encode(), n, _p(pos), _p(ref_code), _p(probs), _p(dp), _p(alt), _p(fwd), _p(rev), 1 if haploid else 0, int(threads),
            _p(out), out.size, _p(off), _p(ok))
    if tot < 0:
        raise NcError(int(tot), "nc_format_snp_records failed")
    return out[:tot].tobytes(), off, ok[:n].astype(bool)


class Context:
    """One GPU, one stream — the analogue of one reference worker process."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = ctypes.c_void_p()
        rc = self._lib.nc_create(device, ctypes.byref(h))
        if rc != NC_OK:
            raise NcError(rc, "nc_create(device=%d) failed: no usable sm_100 CUDA device (there is no CPU fallback)" % device)
        self._h = h
        self.device = device
        self.n_sites = 0
        self.n_isites = 0
        self._indel_haploid = False
        self.n_chunks = 0
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != NC_OK:
            raise NcError(rc, (self._lib.nc_last_error(self._h) or b"").decode())

    # ---- staging
    def stage_reads(self, rs, ref_start=0):
        arrs = (rs.pos, rs.flag, rs.cigar_off, rs.cigar, rs.seq_off, rs.l_seq, rs.seq4, rs.ref)
        self._keep = arrs
        self._check(self._lib.nc_stage_reads(self._h, rs.n, _p(rs.pos), _p(rs.flag), _p(rs.cigar_off), _p(rs.cigar),
                                             _p(rs.seq_off), _p(rs.l_seq), _p(rs.seq4), _p(rs.ref), ref_start, len(rs.ref)))

    def stage_arrays(self, pos, flag, cigar_off, cigar, seq_off, l_seq, seq4, ref, ref_start=0):
        self._keep = (pos, flag, cigar_off, cigar, seq_off, l_seq, seq4, ref)
        self._check(self._lib.nc_stage_reads(self._h, len(pos), _p(pos), _p(flag), _p(cigar_off), _p(cigar), _p(seq_off),
                                             _p(l_seq), _p(seq4), _p(ref), ref_start, len(ref)))

    def stage_tags(self, hp, ps):
        hp = np.ascontiguousarray(hp, np.int8)
        ps = np.ascontiguousarray(ps, np.int32)
        self._keep_tags = (hp, ps)
        self._check(self._lib.nc_stage_tags(self._h, _p(hp), _p(ps)))

    # ---- device-side BAM input
    def bam_device_open(self, path):
        """Inflate + parse `path` on the GPU -> list of (name, length, n_reads, n_tagged) in header order."""
        n = ctypes.c_int32(0)
        self._check(self._lib.nc_bam_device_open(self._h, os.fsencode(path), ctypes.byref(n)))
        out = []
        for i in range(n.value):
            c = NcBamDeviceContig()
            self._check(self._lib.nc_bam_device_contig(self._h, i, ctypes.byref(c)))
            out.append((c.name.decode(), int(c.length), int(c.n_reads), int(c.n_tagged)))
        return out

    def bam_device_stage(self, index, ref, ref_start=0):
        ref = np.ascontiguousarray(ref, np.uint8)
        self._keep = (ref,)
        self._check(self._lib.nc_bam_device_stage(self._h, int(index), _p(ref), int(ref_start), len(ref)))

    def bam_device_close(self):
        self._check(self._lib.nc_bam_device_close(self._h))

    def bam_device_timings(self):
        ms = (ctypes.c_float * 4)()
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        self._check(self._lib.nc_bam_device_timings(self._h, ctypes.byref(ms), ctypes.byref(a), ctypes.byref(b)))
        return {"host_ms": ms[0], "h2d_ms": ms[1], "inflate_ms": ms[2], "records_ms": ms[3], "compressed_bytes": a.value, "inflated_bytes": b.value,
                "record_walk": "parallel from the BAI's record starts" if self._lib.nc_bam_device_walk_mode(self._h) == 1 else "single walker"}

    def fetch_staged(self):
        """The staged contig back on the host (validation): dict of the BAM-native arrays."""
        lib = self._lib
        i64 = ctypes.c_int64
        lib.nc_debug_fetch_staged.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(i64)] * 3 + [ctypes.c_void_p] * 9
        lib.nc_debug_fetch_staged.restype = ctypes.c_int
        n, nc_, ns = i64(0), i64(0), i64(0)
        self._check(lib.nc_debug_fetch_staged(self._h, ctypes.byref(n), ctypes.byref(nc_), ctypes.byref(ns), *([None] * 9)))
        d = dict(pos=np.empty(n.value, np.int32), flag=np.empty(n.value, np.uint16), cigar_off=np.empty(n.value + 1, np.int64),
                 cigar=np.empty(nc_.value, np.uint32), seq_off=np.empty(n.value + 1, np.int64), l_seq=np.empty(n.value, np.int32),
                 seq4=np.empty(ns.value, np.uint8), hp=np.zeros(n.value, np.int8), ps=np.zeros(n.value, np.int32))
        self._check(lib.nc_debug_fetch_staged(self._h, None, None, None, _p(d["pos"]), _p(d["flag"]), _p(d["cigar_off"]), _p(d["cigar"]) if nc_.value else None,
                                              _p(d["seq_off"]), _p(d["l_seq"]), _p(d["seq4"]) if ns.value else None, _p(d["hp"]), _p(d["ps"])))
        return d

    # ---- indel feature path
    def indel_scan(self, params, chunks, bed=None):
        ch = np.array([(int(s), int(e)) for s, e in chunks], dtype=CHUNK_DTYPE)
        bd = np.ascontiguousarray(np.array(bed, dtype=np.int32).reshape(-1, 2)) if bed is not None and len(bed) else None
        n = ctypes.c_int64(0)
        self._check(self._lib.nc_indel_scan(self._h, ctypes.byref(params), _p(ch) if len(ch) else None, len(ch), _p(bd),
                                            0 if bd is None else len(bd), ctypes.byref(n)))
        out = np.empty(n.value, VARIANT_DTYPE)
        self._check(self._lib.nc_indel_fetch_variants(self._h, _p(out) if n.value else None))
        return out

    def indel_build(self, params, chunks, sites, want_tensors=True, fetch=True):
        """Pass 2 + msa for `sites`; -> (meta, tensors or None, cns).  want_tensors=False leaves the tensors on the device
        (for `indel_forward`); fetch=False copies nothing back (-> None)."""
        ch = np.array([(int(s), int(e)) for s, e in chunks], dtype=CHUNK_DTYPE)
        sites = np.ascontiguousarray(sites, VARIANT_DTYPE)
        n = len(sites)
        self._check(self._lib.nc_indel_build(self._h, ctypes.byref(params), _p(ch) if len(ch) else None, len(ch), _p(sites) if n else None, n))
        self.n_isites, self._indel_haploid = n, bool(params.haploid)
        if not fetch:
            return None
        meta = np.empty(n, INDEL_META_DTYPE)
        tensors = np.empty((n, 3, 5, 128, 2), np.float32) if want_tensors else None
        cns = np.empty((n, 3, INDEL_CNS_MAX), np.uint8)
        self._check(self._lib.nc_indel_fetch(self._h, _p(meta) if n else None, _p(tensors) if (n and want_tensors) else None, _p(cns) if n else None))
        return meta, tensors, cns

    def indel_fetch(self, want_tensors=False):
        n = self.n_isites
        meta = np.empty(n, INDEL_META_DTYPE)
        tensors = np.empty((n, 3, 5, 128, 2), np.float32) if want_tensors else None
        cns = np.empty((n, 3, INDEL_CNS_MAX), np.uint8)
        self._check(self._lib.nc_indel_fetch(self._h, _p(meta) if n else None, _p(tensors) if (n and want_tensors) else None, _p(cns) if n else None))
        return meta, tensors, cns

    def indel_fetch_alleles(self):
        """Device allele prediction of the last build: int32 [n_sites, 3, 2] = (ref allele length, alt allele length), -1 = none."""
        out = np.empty((self.n_isites, 3, 2), np.int32)
        self._check(self._lib.nc_indel_fetch_alleles(self._h, _p(out) if self.n_isites else None))
        return out

    def indel_forward(self, impl=0, fetch=True):
        """Indel CNN on the device-resident tensors of the last `indel_build` -> float32 [n_sites, 4] (haploid: [n_sites, 1])."""
        n = self.n_isites
        probs = np.empty((n, 1 if self._indel_haploid else 4), np.float32) if fetch else None
        self._check(self._lib.nc_indel_forward(self._h, impl, _p(probs) if fetch and n else None))
        return probs

    def indel_fetch_probs(self, out):
        """Probabilities of the last `indel_forward` into a caller-owned (e.g. pinned) float32 array [n_sites, 4 or 1]."""
        if self.n_isites == 0:
            return
        assert out.nbytes >= self.n_isites * (4 if self._indel_haploid else 16)
        self._check(self._lib.nc_indel_fetch_probs(self._h, _p(out)))

    def indel_timings(self):
        t = NcIndelTimings()
        self._check(self._lib.nc_get_indel_timings(self._h, ctypes.byref(t)))
        return {k: getattr(t, k) for k, _ in NcIndelTimings._fields_ if k != "reserved"}

    def decode_reads(self):
        self._check(self._lib.nc_decode_reads(self._h))

    # ---- SNP feature path
    def snp_scan(self, params, chunks, bed=None):
        """chunks: iterable of (start, end) 1-based inclusive; bed: iterable of (start, end) or None."""
        ch = np.array([(int(s), int(e)) for s, e in chunks], dtype=CHUNK_DTYPE)
        bd = np.ascontiguousarray(np.array(bed, dtype=np.int32).reshape(-1, 2)) if bed is not None and len(bed) else None
        n = ctypes.c_int64(0)
        self._check(self._lib.nc_snp_scan(self._h, ctypes.byref(params), _p(ch) if len(ch) else None, len(ch), _p(bd),
                                          0 if bd is None else len(bd), ctypes.byref(n)))
        self.n_sites, self.n_chunks = n.value, len(ch)
        return n.value

    def snp_fetch(self, want_mat=True):
        n, nc = self.n_sites, self.n_chunks
        mat = np.empty((n, SITE_STRIDE), np.int16) if want_mat else None
        meta = np.empty(n, META_DTYPE)
        depth = np.zeros(nc, np.float64)
        count = np.zeros(nc, np.int64)
        self._check(self._lib.nc_snp_fetch(self._h, _p(mat) if want_mat and n else None, _p(meta) if n else None,
                                           _p(depth) if nc else None, _p(count) if nc else None))
        return mat, meta, depth, count

    def snp_fetch_range(self, first, count):
        mat = np.empty((count, SITE_STRIDE), np.int16)
        self._check(self._lib.nc_snp_fetch_range(self._h, int(first), int(count), _p(mat) if count else None))
        return mat

    def indel_fetch_range(self, first, count):
        t = np.empty((count, 3, 5, 128, 2), np.float32)
        self._check(self._lib.nc_indel_fetch_range(self._h, int(first), int(count), _p(t) if count else None))
        return t

    # ---- models
    def load_snp_weights(self, blob, train_coverage, haploid):
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        self._check(self._lib.nc_load_snp_weights(self._h, _p(blob), blob.size, float(train_coverage), 1 if haploid else 0))

    def load_indel_weights(self, blob, haploid):
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        self._check(self._lib.nc_load_indel_weights(self._h, _p(blob), blob.size, 1 if haploid else 0))

    def snp_forward(self, normalize=True, impl=0, fetch=True):
        probs = np.empty((self.n_sites, 4), np.float32) if fetch else None
        self._check(self._lib.nc_snp_forward(self._h, 1 if normalize else 0, impl, _p(probs) if fetch and self.n_sites else None))
        return probs

    def fetch_calls(self, probs, meta):
        """Per-site call records of the last scan + forward into caller-owned (e.g. pinned) arrays:
        probs float32 [n_sites,4], meta META_DTYPE-sized rows [n_sites]."""
        if self.n_sites == 0:
            return
        assert probs.nbytes >= self.n_sites * 16 and meta.nbytes >= self.n_sites * META_DTYPE.itemsize
        self._check(self._lib.nc_snp_fetch(self._h, None, _p(meta), None, None))
        self._check(self._lib.nc_snp_fetch_probs(self._h, _p(probs)))

    def snp_model_forward(self, x, ref_onehot, haploid=False, impl=0):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 5, 41, 5)
        ref = np.ascontiguousarray(ref_onehot, dtype=np.float32).reshape(-1, 4)
        n = len(x)
        out = np.empty((n, 4 if haploid else 10), np.float32)
        self._check(self._lib.nc_snp_model_forward(self._h, _p(x), _p(ref), n, 1 if haploid else 0, impl, _p(out)))
        return out

    def indel_model_forward(self, x, haploid=False, impl=0):
        rows = 5 if haploid else 15
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, rows, 128, 2)
        n = len(x)
        out = np.empty((n, 1 if haploid else 4), np.float32)
        self._check(self._lib.nc_indel_model_forward(self._h, _p(x), n, 1 if haploid else 0, impl, _p(out)))
        return out

    def sync(self):
        self._check(self._lib.nc_sync(self._h))

    def set_blocking_sync(self, on=True):
        """Host waits sleep instead of spinning (several contexts / ranks on few host cores)."""
        self._check(self._lib.nc_set_blocking_sync(self._h, 1 if on else 0))

    def timings(self):
        t = NcTimings()
        self._check(self._lib.nc_get_timings(self._h, ctypes.byref(t)))
        return {k: getattr(t, k) for k, _ in NcTimings._fields_}

    def event_record(self, slot):
        self._check(self._lib.nc_event_record(self._h, slot))

    def event_elapsed_ms(self, a, b):
        ms = ctypes.c_float()
        self._check(self._lib.nc_event_elapsed_ms(self._h, a, b, ctypes.byref(ms)))
        return ms.value

    def invalidate_decode(self):
        self._check(self._lib.nc_invalidate_decode(self._h))

    def sm_count(self):
        return self._lib.nc_device_sm_count(self._h)

    def device_buffers(self):
        m, me, pr = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        n = ctypes.c_int64()
        self._check(self._lib.nc_snp_device_buffers(self._h, ctypes.byref(m), ctypes.byref(me), ctypes.byref(pr), ctypes.byref(n)))
        return m.value, me.value, pr.value, n.value
