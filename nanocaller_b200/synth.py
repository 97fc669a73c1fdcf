"""Python face of the synthetic world generator (csrc/synth.cpp): seeded reference + truth + reads.

Presets follow SURVEY.md §8(d): `ont` (log-normal 12 kb reads, 3/2/1 % sub/del/ins),
`hifi` (N(15 kb, 2 kb)-like, 0.1/0.05/0.05 %).  No network, no files: arrays only.
"""
import ctypes
import os

import numpy as np

from . import build as _build
from .host.readset import ReadSet


class _Params(ctypes.Structure):
    _fields_ = [
        ("seed", ctypes.c_uint64), ("contig_len", ctypes.c_int64), ("coverage", ctypes.c_double),
        ("len_median", ctypes.c_double), ("len_sigma", ctypes.c_double),
        ("len_min", ctypes.c_int32), ("len_max", ctypes.c_int32),
        ("sub_rate", ctypes.c_double), ("del_rate", ctypes.c_double), ("ins_rate", ctypes.c_double),
        ("het_every", ctypes.c_int32), ("hom_every", ctypes.c_int32), ("sys_per_10k", ctypes.c_int32),
        ("clip_prob", ctypes.c_double), ("clip_max", ctypes.c_int32), ("ploidy", ctypes.c_int32),
        ("mask_every", ctypes.c_int32), ("mask_len", ctypes.c_int32),
        ("junk_frac", ctypes.c_double), ("nbase_rate", ctypes.c_double),
        ("indel_every", ctypes.c_int32), ("indel_maxlen", ctypes.c_int32),
        ("untagged_frac", ctypes.c_double),
    ]


PRESETS = {
    "ont": dict(len_median=12000.0, len_sigma=0.6, len_min=1000, len_max=100000,
                sub_rate=0.03, del_rate=0.02, ins_rate=0.01),
    "hifi": dict(len_median=15000.0, len_sigma=0.13, len_min=5000, len_max=30000,
                 sub_rate=0.001, del_rate=0.0005, ins_rate=0.0005),
    "short_ont": dict(len_median=4000.0, len_sigma=0.5, len_min=500, len_max=20000,
                      sub_rate=0.03, del_rate=0.02, ins_rate=0.01),
}

DEFAULTS = dict(seed=20, contig_len=1_000_000, coverage=30.0, het_every=1000, hom_every=3000,
                sys_per_10k=100, clip_prob=0.2, clip_max=50, ploidy=2, mask_every=0, mask_len=0,
                junk_frac=0.0, nbase_rate=0.0, indel_every=0, indel_maxlen=0, untagged_frac=0.0)

_lib = None


def _load():
    global _lib
    if _lib is None:
        path = _build.build_synth()
        lib = ctypes.CDLL(path)
        P = ctypes.POINTER(_Params)
        vp = ctypes.c_void_p
        lib.nc_synth_world.argtypes = [P, vp, vp, vp, ctypes.c_int]
        lib.nc_synth_num_reads.argtypes = [P]
        lib.nc_synth_num_reads.restype = ctypes.c_int64
        lib.nc_synth_count.argtypes = [P, vp, vp, vp, ctypes.c_int64, vp, vp, vp, vp, vp, vp, ctypes.c_int]
        lib.nc_synth_fill.argtypes = [P, vp, vp, vp, ctypes.c_int64, vp, vp, vp, vp, ctypes.c_int]
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class World:
    """A generated contig: `.reads` (ReadSet), `.var` / `.indel` truth bytes, `.params`."""

    def __init__(self, reads, var, indel, params):
        self.reads, self.var, self.indel, self.params = reads, var, indel, params

    def truth_snps(self):
        """(pos1, kind, alt_char) for het(1)/hom(2) truth SNPs; positions 1-based."""
        idx = np.nonzero((self.var & 3) % 3 != 0)[0]
        kinds = self.var[idx] & 3
        alts = np.frombuffer(b"ACGT", np.uint8)[(self.var[idx] >> 2) & 3]
        return idx + 1, kinds, alts


def make_world(chrom="chr20", preset="ont", nthreads=None, **kw):
    lib = _load()
    cfg = dict(DEFAULTS)
    cfg.update(PRESETS[preset])
    cfg.update(kw)
    p = _Params(**cfg)
    if nthreads is None:
        nthreads = min(32, os.cpu_count() or 1)
    L = int(cfg["contig_len"])
    ref = np.empty(L, np.uint8)
    var = np.empty(L, np.uint8)
    indel = np.zeros(L, np.uint8)
    lib.nc_synth_world(ctypes.byref(p), _ptr(ref), _ptr(var), _ptr(indel), nthreads)
    use_indel = cfg["indel_every"] > 0
    ind_ptr = _ptr(indel) if use_indel else None
    n = lib.nc_synth_num_reads(ctypes.byref(p))
    pos = np.empty(n, np.int32)
    ncig = np.empty(n, np.int32)
    lseq = np.empty(n, np.int32)
    flag = np.empty(n, np.uint16)
    hap = np.empty(n, np.int8)
    span = np.empty(n, np.int32)
    lib.nc_synth_count(ctypes.byref(p), _ptr(ref), _ptr(var), ind_ptr, n, _ptr(pos), _ptr(ncig), _ptr(lseq),
                       _ptr(flag), _ptr(hap), _ptr(span), nthreads)
    cig_off = np.zeros(n + 1, np.int64)
    np.cumsum(ncig, out=cig_off[1:])
    seq_off = np.zeros(n + 1, np.int64)
    np.cumsum((lseq.astype(np.int64) + 1) // 2, out=seq_off[1:])
    cigar = np.empty(cig_off[-1], np.uint32)
    seq4 = np.zeros(seq_off[-1], np.uint8)
    lib.nc_synth_fill(ctypes.byref(p), _ptr(ref), _ptr(var), ind_ptr, n, _ptr(cig_off), _ptr(seq_off),
                      _ptr(cigar), _ptr(seq4), nthreads)
    ps = np.where(hap > 0, 1, 0).astype(np.int32)
    rs = ReadSet(chrom, ref, pos, flag, cig_off, cigar, seq_off, lseq, seq4, hp=hap, ps=ps)
    rs._ref_end = (pos.astype(np.int64) + span).astype(np.int32)
    return World(rs, var, indel, cfg)
