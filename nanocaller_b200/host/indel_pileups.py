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
    last imputed hit on a key and wins in pass 2 (:309).  Device hits arrive in (chunk, column) order per chunk."""
    out = []
    for c in np.unique(variants["chunk"]):
        sel = variants[variants["chunk"] == c]
        d, extra = {}, {}
        for k, t, src in zip(sel["key"].tolist(), sel["type"].tolist(), sel["src"].tolist()):
            d[k] = t
            if src:
                extra[k] = src
        for k in sorted(d):
            out.append((k, d[k], int(c), extra.get(k, 0)))
    return np.array(out, dtype=capi.VARIANT_DTYPE) if out else np.zeros(0, capi.VARIANT_DTYPE)


def candidates_for_chunks(ctx, rs, dct, chunks, bed=None, haploid=False):
    """Scan + build for a list of chunk dicts of one contig; -> per-chunk reference-shaped tuples
    (diploid: 6-tuple of generate_indel_pileups.py:370; haploid: 3-tuple of generate_indel_pileups_haploid.py:277)."""
    snp_pileups.stage(ctx, rs)
    ctx.stage_tags(rs.hp, rs.ps)
    P = capi.indel_params(dct, haploid)
    ch = [(c["start"], c["end"]) for c in chunks]
    variants = ctx.indel_scan(P, ch, bed)
    # hits of one chunk are produced in column order by one warp; keep that order when applying the dict semantics
    sites = order_variants(variants)
    meta, tensors, cns = ctx.indel_build(P, ch, sites)
    max_range = {0: max(10, int(dct["win_size"])), 1: 10}
    # ---- allele prediction for every kept site and read group in one batched, multithreaded library call
    okmask = meta["ok"][:, 2] > 0 if haploid else meta["ok"].min(1) > 0
    kept = np.nonzero(okmask)[0]
    groups = (2,) if haploid else (0, 1, 2)
    pred = {}
    if len(kept):
        cap = cns.shape[2]
        ref_rows, ref_off, ref_len = [], [], []
        off = 0
        for s in kept:
            p0 = int(meta["pos"][s]) - 1
            m = int(meta["ref_len"][s])
            ref_rows.append(rs.ref[p0:p0 + m])
            ref_off.append(off); ref_len.append(m)
            off += m
        ref_bytes = np.concatenate(ref_rows)
        ref_flat = _REF_CODE[ref_bytes]
        it_site = np.repeat(kept, len(groups))
        it_grp = np.tile(np.asarray(groups), len(kept))
        alt_off = (it_site.astype(np.int64) * cns.shape[1] + it_grp) * cap
        alt_len = meta["cns_len"][it_site, it_grp].astype(np.int32)
        r_off = np.repeat(np.asarray(ref_off, np.int64), len(groups))
        r_len = np.repeat(np.asarray(ref_len, np.int32), len(groups))
        mr = np.array([max_range[int(t)] for t in meta["type"][it_site]], np.int32)
        ro, ao = capi.allele_predict_batch(cns.reshape(-1), alt_off, alt_len, ref_flat, r_off, r_len, mr)
        ref_txt = ref_bytes.tobytes().decode()
        for k in range(len(it_site)):
            s, g = int(it_site[k]), int(it_grp[k])
            if ro[k] < 0:
                pred[(s, g)] = (None, None)
            else:
                b0 = int(r_off[k])
                alt = BASES[cns[s, g, :alt_len[k]]].tobytes().decode()
                pred[(s, g)] = (ref_txt[b0:b0 + min(int(ro[k]), int(r_len[k]))], alt[:int(ao[k])])
    res = []
    for ci in range(len(chunks)):
        sel = np.nonzero((meta["chunk"] == ci) & okmask)[0]
        if len(sel) == 0:
            res.append(([], [], []) if haploid else ([], [], [], [], [], []))       # generate_indel_pileups.py:363-364
            continue
        pos = [int(p) for p in meta["pos"][sel]]
        alleles, phase = [], []
        for s in sel:
            trip = [pred[(int(s), g)] for g in groups]
            alleles.append(trip[0] if haploid else trip)
            ph = int(meta["phase"][s])
            phase.append(None if ph == -1 else ph)                              # imputed site whose first hap0 read has no HP tag (:355)
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
