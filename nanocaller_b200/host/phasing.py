"""Read-based phasing of the SNP calls and haplotagging of the reads between the SNP and the indel stage.

The reference shells out to WhatsHap here (indelCaller.phase_run, indelCaller.py:190-262): SNP records with
QUAL >= --phase_qual_score are phased (`whatshap phase`, :237), the BAM is haplotagged (`whatshap haplotag`, :244) and the
indel stage reads the HP / PS tags (generate_indel_pileups.py:180-188).  WhatsHap is not part of the reference repository
and is not available, so the algorithm is this package's own (libnc_phase.so, include/nanocaller_b200_phase.h,
DESIGN.md §4.6): same inputs, same outputs (phased GT + PS in the SNP records, HP / PS per read), validated against the
synthetic generator's true haplotypes rather than against WhatsHap.  Tags are set on the in-memory ReadSet — no phased
BAM is written; the indel stage stages them with `nc_stage_tags`."""
import ctypes
import os

import numpy as np

_NIB = {"A": 1, "C": 2, "G": 4, "T": 8}
PHASE_EXPORTS = ["nc_phase_read_alleles", "nc_phase_sites"]
_lib = None


def load_phase_library():
    """ctypes handle of libnc_phase.so (built in-tree by nanocaller_b200.build)."""
    global _lib
    if _lib is None:
        from .. import build
        lib = ctypes.CDLL(build.build_phase())
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
        lib.nc_phase_read_alleles.argtypes = [i64, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i32]
        lib.nc_phase_sites.argtypes = [i64, vp, vp, vp, vp, i64, i32, vp, vp, vp, vp]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def phase_sites(rs, site_pos0, nib_a, nib_b, supplementary=False, iterations=2, threads=0):
    """Phase heterozygous sites of one contig and haplotag its reads.
    site_pos0 ascending 0-based positions; nib_a / nib_b the two alleles as BAM nibbles.
    -> (site_hap int8 [n] (0 A|B, 1 B|A, -1 unphased), site_ps int32 [n] (1-based position of the block's first site, 0 unphased),
        read_hp int8 [n_reads], read_ps int32 [n_reads])."""
    lib = load_phase_library()
    site_pos0 = np.ascontiguousarray(site_pos0, np.int32)
    nib_a = np.ascontiguousarray(nib_a, np.uint8)
    nib_b = np.ascontiguousarray(nib_b, np.uint8)
    n, ns = rs.n, len(site_pos0)
    assert ns == 0 or np.all(np.diff(site_pos0) > 0), "sites must be strictly ascending"
    first = np.searchsorted(site_pos0, rs.pos, side="left").astype(np.int64)
    last = np.searchsorted(site_pos0, rs.ref_end, side="left").astype(np.int64)
    mapped = (rs.flag & 0x4) == 0
    cnt = np.where(mapped, np.maximum(last - first, 0), 0)
    pair_off = np.zeros(n + 1, np.int64)
    np.cumsum(cnt, out=pair_off[1:])
    allele = np.empty(max(1, int(pair_off[-1])), np.uint8)
    rc = lib.nc_phase_read_alleles(n, _p(rs.pos), _p(rs.cigar_off), _p(rs.cigar), _p(rs.seq_off), _p(rs.l_seq), _p(rs.seq4), ns,
                                   _p(site_pos0), _p(nib_a), _p(nib_b), _p(first), _p(pair_off), _p(allele), threads)
    if rc:
        raise RuntimeError("nc_phase_read_alleles failed (%d)" % rc)
    flag_filter = (0x4 | 0x100 | 0x200 | 0x400) if supplementary else (0x4 | 0x100 | 0x200 | 0x400 | 0x800)
    use = np.ascontiguousarray(((rs.flag & flag_filter) == 0).astype(np.uint8))
    site_hap = np.empty(ns, np.int8)
    site_block = np.empty(ns, np.int32)
    read_hp = np.empty(n, np.int8)
    read_block = np.empty(n, np.int32)
    rc = lib.nc_phase_sites(n, _p(use), _p(first), _p(pair_off), _p(allele), ns, iterations, _p(site_hap), _p(site_block), _p(read_hp), _p(read_block))
    if rc:
        raise RuntimeError("nc_phase_sites failed (%d)" % rc)
    ps_of = lambda blk: np.where(blk >= 0, site_pos0[np.maximum(blk, 0)] + 1, 0).astype(np.int32) if ns else np.zeros(len(blk), np.int32)
    return site_hap, ps_of(site_block), read_hp, ps_of(read_block)


def phase_snp_records(lines, rs, phase_qual_score=10.0, supplementary=False, threads=0):
    """`whatshap phase` + `whatshap haplotag` for the records of ONE contig (`rs.chrom`).
    lines: SNP record lines of the PASS file (snpCaller.py record format, FORMAT GT:DP:VF:AD:ADF:ADR).  Heterozygous single-base calls
    (`0/1`, `1/2`) with QUAL >= phase_qual_score (indelCaller.py:232) are phased: GT becomes `0|1` / `1|0` (`1|2` / `2|1`) and PS is
    appended to FORMAT, as WhatsHap writes it; every other line is returned unchanged.  Sets rs.hp / rs.ps.
    -> (new lines, stats dict)."""
    sel, pos0, na, nb = [], [], [], []
    last = -1
    for k, ln in enumerate(lines):
        f = ln.rstrip("\n").split("\t")
        if f[0] != rs.chrom or float(f[5]) < phase_qual_score:
            continue
        gt = f[9].split(":", 1)[0]
        alts = f[4].split(",")
        if gt == "0/1" and len(alts) == 1:
            a, b = f[3], alts[0]
        elif gt == "1/2" and len(alts) == 2:
            a, b = alts
        else:
            continue
        p = int(f[1]) - 1
        if a not in _NIB or b not in _NIB or p <= last:             # single bases only; duplicate positions (shared chunk ends) phase once
            continue
        last = p
        sel.append(k); pos0.append(p); na.append(_NIB[a]); nb.append(_NIB[b])
    site_hap, site_ps, read_hp, read_ps = phase_sites(rs, np.asarray(pos0, np.int32), np.asarray(na, np.uint8), np.asarray(nb, np.uint8),
                                                     supplementary=supplementary, threads=threads)
    out = list(lines)
    for k, hj, ps in zip(sel, site_hap.tolist(), site_ps.tolist()):
        if hj < 0:
            continue
        f = out[k].rstrip("\n").split("\t")
        gt, rest = f[9].split(":", 1)
        x, y = gt.split("/")
        f[8] += ":PS"
        f[9] = "%s|%s:%s:%d" % ((x, y, rest, ps) if hj == 0 else (y, x, rest, ps))       # hj = 1: haplotype 1 carries the second allele
        out[k] = "\t".join(f) + "\n"
    rs.hp = np.ascontiguousarray(read_hp, np.int8)
    rs.ps = np.ascontiguousarray(read_ps, np.int32)
    stats = {"het_sites": len(sel), "phased_sites": int((site_hap >= 0).sum()), "blocks": int(len(np.unique(site_ps[site_ps > 0]))),
             "tagged_reads": int((read_hp > 0).sum()), "reads": int(rs.n)}
    return out, stats
