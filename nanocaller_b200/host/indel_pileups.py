"""Host mirror of nanocaller_src/generate_indel_pileups.py (diploid) over the CUDA library.

`get_indel_testing_candidates(dct, chunk)` keeps the reference signature and return tuple
(generate_indel_pileups.py:129, :370; call site indelCaller.py:69).  Candidate scan, read slices, alignment, tensors
and consensus are computed in libnanocaller_b200.so; this module applies the dict semantics of `variants`, turns the
consensus / reference pair into allele strings (allele_prediction, :77-127, alignment by the library's nc_nw_trace) and
reshapes the results.  MUSCLE and parasail are replaced by the library's own alignments."""
import numpy as np

from . import capi, snp_pileups, sources

BASES = np.frombuffer(b"AGTC", np.uint8)
_REF_CODE = np.full(256, 4, np.uint8)
for _i, _c in enumerate(b"AGTC"):
    _REF_CODE[_c] = _i


def allele_prediction(alt, ref_seq, max_range, alt_codes=None, ref_codes=None):
    """generate_indel_pileups.py:77-127, control flow kept line for line; the CIGAR comes from nc_nw_trace."""
    if alt_codes is None:
        alt_codes = _REF_CODE[np.frombuffer(alt.encode(), np.uint8)]
    if ref_codes is None:
        ref_codes = _REF_CODE[np.frombuffer(ref_seq.encode(), np.uint8)]
    cigar_op = capi.nw_trace(alt_codes, ref_codes, 9, 1, 20, -10)
    indel = False
    ref_cnt = [0] * 10
    alt_cnt = [0] * 10
    mis_match_cnt_before_indel = False
    mis_match_cnt_after_indel = (0, 0)
    for op, cnt in cigar_op:
        if op == 8 or op == 7:
            ref_cnt[op] += cnt
            alt_cnt[op] += cnt
            if indel:
                mis_match_cnt_after_indel[op - 7] += cnt
            else:
                mis_match_cnt_before_indel = True
        if op == 1:
            alt_cnt[op] += cnt
            mis_match_cnt_after_indel = [0, 0]
            indel = True
        if op == 2:
            ref_cnt[op] += cnt
            mis_match_cnt_after_indel = [0, 0]
            indel = True
        if indel is False and sum(ref_cnt) >= max_range + 10:
            if ref_cnt[8]:
                out_len = sum(ref_cnt) if op == 8 else sum(ref_cnt) - cnt
                return ref_seq[:out_len], alt[:out_len]
            else:
                return (None, None)
        if indel is True:
            if sum(mis_match_cnt_after_indel) > 20:
                break
    ref_out_len = sum(ref_cnt) if op == 8 else sum(ref_cnt) - cnt
    alt_out_len = sum(alt_cnt) if op == 8 else sum(alt_cnt) - cnt
    if not mis_match_cnt_before_indel:
        ref_out_len += 1
        alt_out_len += 1
    return ref_seq[:ref_out_len], alt[:alt_out_len]


def order_variants(variants):
    """The reference keeps `variants` in a dict (a later hit on the same key overwrites the type, :268,:274,:301) and visits
    the keys in column order in pass 2 (:306-320); `extra_variants` (:302) is a second dict that keeps the source column of the
    last imputed hit on a key and wins in pass 2 (:309).  Device hits arrive in column order within a chunk (one warp per chunk
    appends them); chunks interleave.  Vectorised: stable sort by (chunk, key), the last entry of every group gives the type,
    the last entry with a source column gives `src`."""
    n = len(variants)
    if n == 0:
        return np.zeros(0, capi.VARIANT_DTYPE)
    key, chunk = variants["key"].astype(np.int64), variants["chunk"].astype(np.int64)
    order = np.lexsort((np.arange(n), key, chunk))
    k, c, t, src = key[order], chunk[order], variants["type"][order], variants["src"][order]
    first = np.ones(n, bool)
    first[1:] = (k[1:] != k[:-1]) | (c[1:] != c[:-1])
    last = np.ones(n, bool)
    last[:-1] = first[1:]
    ar = np.arange(n)
    start = np.maximum.accumulate(np.where(first, ar, 0))                 # index of the group's first entry
    nz = np.maximum.accumulate(np.where(src != 0, ar, -1))                # index of the latest entry with a source column
    src_last = np.where(nz >= start, src[np.maximum(nz, 0)], 0)
    out = np.zeros(int(last.sum()), capi.VARIANT_DTYPE)
    out["key"], out["type"], out["chunk"], out["src"] = k[last], t[last], c[last], src_last[last]
    return out


def scan_build(ctx, rs, dct, chunks, bed=None, haploid=False, want_tensors=True):
    """Pass 1 (scan), the dict semantics of `variants`, pass 2 + msa (build) for a list of chunk dicts of one contig.
    -> (meta, tensors or None, cns); with want_tensors=False the tensors stay on the device for `Context.indel_forward`."""
    snp_pileups.stage(ctx, rs)
    if not isinstance(rs, sources.DeviceContig):          # a device-decoded contig brings its HP / PS tags along
        ctx.stage_tags(rs.hp, rs.ps)
    P = capi.indel_params(dct, haploid)
    ch = [(c["start"], c["end"]) for c in chunks]
    variants = ctx.indel_scan(P, ch, bed)
    sites = order_variants(variants)
    return ctx.indel_build(P, ch, sites, want_tensors=want_tensors)


def kept_sites(meta, haploid):
    """msa() succeeded for every group the caller uses (generate_indel_pileups.py:342-348)."""
    return meta["ok"][:, 2] > 0 if haploid else meta["ok"].min(1) > 0


class AllelePredictions:
    """allele_prediction (generate_indel_pileups.py:77-127) for every kept site and read group, as arrays: item k is site
    `site[k]`, group `grp[k]`; ref_out / alt_out = lengths of the allele strings (-1 / -1: the reference returns (None, None))."""

    def __init__(self, rs, dct, meta, cns, haploid, threads=0, device_lengths=None):
        """device_lengths: int32 [n_sites, 3, 2] from nc_indel_fetch_alleles (the alignment ran on the GPU right after msa); without it,
        or for items it marks -2, the same alignment runs on host threads (nc_allele_predict_batch).  The reference windows are only
        gathered on the host when something needs them (host alignment, allele strings)."""
        self.groups = (2,) if haploid else (0, 1, 2)
        self.cns, self.meta, self.rs = cns, meta, rs
        kept = np.nonzero(kept_sites(meta, haploid))[0]
        self.kept = kept
        ng = len(self.groups)
        self.site = np.repeat(kept, ng)
        self.grp = np.tile(np.asarray(self.groups), len(kept))
        self._windows = None
        self.alt_len = meta["cns_len"][self.site, self.grp].astype(np.int32) if len(kept) else np.zeros(0, np.int32)
        if device_lengths is not None:
            dl = np.asarray(device_lengths, np.int32)[self.site, self.grp]
            self.ref_out, self.alt_out = dl[:, 0].copy(), dl[:, 1].copy()
            todo = np.nonzero(self.ref_out == -2)[0]
        else:
            self.ref_out, self.alt_out = np.empty(len(self.site), np.int32), np.empty(len(self.site), np.int32)
            todo = np.arange(len(self.site))
        if len(todo):
            ref_bytes, r_off, r_len = self.windows()
            cap = cns.shape[2]
            alt_off = (self.site.astype(np.int64) * cns.shape[1] + self.grp) * cap
            win = max(10, int(dct["win_size"]))
            mr = np.where(meta["type"][self.site] == 0, win, 10).astype(np.int32)   # max_range (:209)
            ro, ao = capi.allele_predict_batch(cns.reshape(-1), alt_off[todo], self.alt_len[todo], _REF_CODE[ref_bytes], r_off[todo], r_len[todo],
                                               mr[todo], threads=threads)
            self.ref_out[todo], self.alt_out[todo] = ro, ao

    def windows(self):
        """(reference bytes of all kept sites' windows back to back, per item offset, per item length)"""
        if self._windows is None:
            kept, ng = self.kept, len(self.groups)
            if len(kept) == 0:
                self._windows = (np.zeros(0, np.uint8), np.zeros(0, np.int64), np.zeros(0, np.int32))
            else:
                p0 = self.meta["pos"][kept].astype(np.int64) - 1
                m = self.meta["ref_len"][kept].astype(np.int64)
                off = np.zeros(len(kept) + 1, np.int64)
                np.cumsum(m, out=off[1:])
                idx = np.repeat(p0 - off[:-1], m) + np.arange(off[-1])             # flat gather of the reference windows
                self._windows = (self.rs.ref[idx], np.repeat(off[:-1], ng), np.repeat(m.astype(np.int32), ng))
        return self._windows

    def strings(self):
        """{(site, group): (ref, alt) or (None, None)}"""
        pred = {}
        ref_bytes, r_off, r_len = self.windows()
        ref_txt = ref_bytes.tobytes().decode()
        cns = self.cns
        for k in range(len(self.site)):
            s, g = int(self.site[k]), int(self.grp[k])
            if self.ref_out[k] < 0:
                pred[(s, g)] = (None, None)
            else:
                b0 = int(r_off[k])
                alt = BASES[cns[s, g, :self.alt_len[k]]].tobytes().decode()
                pred[(s, g)] = (ref_txt[b0:b0 + min(int(self.ref_out[k]), int(r_len[k]))], alt[:int(self.alt_out[k])])
        return pred


def per_chunk_calls(meta, pred, n_chunks, haploid):
    """-> per chunk (site indices, pos list, alleles list, phase list) of the kept sites, in key order."""
    okmask = kept_sites(meta, haploid)
    groups = (2,) if haploid else (0, 1, 2)
    out = []
    for ci in range(n_chunks):
        sel = np.nonzero((meta["chunk"] == ci) & okmask)[0]
        pos = [int(p) for p in meta["pos"][sel]]
        alleles, phase = [], []
        for s in sel:
            trip = [pred[(int(s), g)] for g in groups]
            alleles.append(trip[0] if haploid else trip)
            ph = int(meta["phase"][s])
            phase.append(None if ph == -1 else ph)                              # imputed site whose first hap0 read has no HP tag (:355)
        out.append((sel, pos, alleles, phase))
    return out


def candidates_for_chunks(ctx, rs, dct, chunks, bed=None, haploid=False):
    """Scan + build for a list of chunk dicts of one contig; -> per-chunk reference-shaped tuples
    (diploid: 6-tuple of generate_indel_pileups.py:370; haploid: 3-tuple of generate_indel_pileups_haploid.py:277)."""
    meta, tensors, cns = scan_build(ctx, rs, dct, chunks, bed, haploid)
    pred = AllelePredictions(rs, dct, meta, cns, haploid, device_lengths=ctx.indel_fetch_alleles()).strings()
    res = []
    for sel, pos, alleles, phase in per_chunk_calls(meta, pred, len(chunks), haploid):
        if len(sel) == 0:
            res.append(([], [], []) if haploid else ([], [], [], [], [], []))       # generate_indel_pileups.py:363-364
            continue
        x = tensors[sel].astype(np.float64)                                  # float32 values in a float64 container (:69-71)
        res.append((pos, x[:, 2], alleles) if haploid else (pos, x[:, 0], x[:, 1], x[:, 2], alleles, phase))
    return res


def get_indel_testing_candidates(dct, chunk, device=0):
    """Drop-in for generate_indel_pileups.get_indel_testing_candidates (same dct / chunk keys, same 6-tuple)."""
    ctx = snp_pileups.context(device)
    rs = sources.resolve(chunk["sam_path"], chunk["chrom"])
    bed = sources.bed_intervals(dct.get("exclude_bed"), chunk["chrom"])
    return candidates_for_chunks(ctx, rs, dct, [chunk], bed)[0]


def get_indel_testing_candidates_haploid(dct, chunk, device=0):
    """Drop-in for generate_indel_pileups_haploid.get_indel_testing_candidates_haploid (indelCaller.py:159): 3-tuple."""
    ctx = snp_pileups.context(device)
    rs = sources.resolve(chunk["sam_path"], chunk["chrom"])
    bed = sources.bed_intervals(dct.get("exclude_bed"), chunk["chrom"])
    return candidates_for_chunks(ctx, rs, dct, [chunk], bed, haploid=True)[0]
