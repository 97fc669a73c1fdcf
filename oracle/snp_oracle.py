"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (numpy) restatement of the reference's SNP feature path, reading an in-memory `ReadSet`
instead of pysam.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.

Follows, function by function (paths under /root/reference/nanocaller_src/):
  get_cnd_pos                 generate_SNP_pileups.py:6-101
  get_snp_testing_candidates  generate_SNP_pileups.py:103-278
  get_chunks                  utils.py:67-83
  scale_counts (N1)           snpCaller.py:90-96, 167-173

Pinned: tests/test_oracle_golden.py checks every function here against fixtures produced by the
UNMODIFIED reference modules imported over oracle/shim (tests/golden/make_golden.py).  The pileup
column semantics below the reference (htslib) are unpinned — see oracle/shim/pysam.py.
"""
import numpy as np

_REF_CONSUME = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1] + [0] * 7, dtype=np.int64)
_QRY_CONSUME = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1] + [0] * 7, dtype=np.int64)
# BAM nibble -> reference base code (generate_SNP_pileups.py:104: A0 G1 T2 C3, '*'/'N' 4).
# IUPAC ambiguity codes would raise KeyError in the reference; they are folded into 4 here.
NIB_TO_CODE = np.full(16, 4, np.int8)
NIB_TO_CODE[1], NIB_TO_CODE[4], NIB_TO_CODE[8], NIB_TO_CODE[2] = 0, 1, 2, 3
# ASCII -> code for the reference string: only UPPER-case AGTC count (generate_SNP_pileups.py:137)
REF_TO_CODE = np.full(256, 4, np.int8)
REF_TO_CODE[ord("A")], REF_TO_CODE[ord("G")], REF_TO_CODE[ord("T")], REF_TO_CODE[ord("C")] = 0, 1, 2, 3

# Distance bins of get_cnd_pos as (a, b, k, mode): neighbours p with a < |p - v| <= b, keep k of them,
# mode 'near' = the k closest to v, 'far' = the k farthest inside the bin.  The outermost bin is
# bounded by the strict radius test `abs(cnd_pos - v_pos) < R`, hence b = R - 1.
CND_BINS = {
    "ont": [(0, 2000, 2, "far"), (2000, 5000, 3, "near"), (5000, 10000, 4, "near"),
            (10000, 20000, 5, "near"), (20000, 49999, 6, "near")],                       # :7-23
    "short_ont": [(0, 2000, 5, "near"), (2000, 5000, 10, "near"), (5000, 49999, 5, "near")],  # :25-39
    "ul_ont": [(0, 2000, 2, "far"), (2000, 5000, 2, "near"), (5000, 10000, 3, "near"),
               (10000, 20000, 3, "near"), (20000, 40000, 4, "near"), (40000, 50000, 3, "near"),
               (50000, 99999, 3, "near")],                                               # :41-61
    "ul_ont_extreme": [(0, 10000, 2, "far"), (10000, 20000, 2, "near"), (20000, 50000, 3, "near"),
                       (50000, 75000, 3, "near"), (75000, 100000, 4, "near"),
                       (100000, 200000, 4, "near"), (200000, 299999, 2, "near")],        # :63-83
    "pacbio": [(0, 2000, 4, "far"), (2000, 5000, 5, "near"), (5000, 10000, 5, "near"),
               (10000, 19999, 6, "near")],                                               # :85-99
}


def get_chunks(regions_list, cpu, max_chunk_size=500000, min_chunk_size=10000):
    """utils.py:67-83 — inclusive, shared chunk ends."""
    chunks_list = []
    total_bases = sum(region[2] - region[1] + 1 for region in regions_list)
    chunksize = min(max_chunk_size, max(min_chunk_size, total_bases // cpu + 1))
    for contig, start, end, ploidy in regions_list:
        for chunk in range(start, end, chunksize):
            chunks_list.append({"chrom": contig, "start": chunk, "end": min(end, chunk + chunksize), "ploidy": ploidy})
    return chunks_list


def get_cnd_pos(v_pos, cnd_pos, seq="ont"):
    """generate_SNP_pileups.py:6-101 on a sorted int array of neighbour positions."""
    cnd_pos = np.asarray(cnd_pos, dtype=np.int64)
    left, right = [], []
    for a, b, k, mode in CND_BINS[seq]:
        # left bin: v-b <= p < v-a ; ascending order, 'near' = last k, 'far' = first k
        lo = np.searchsorted(cnd_pos, v_pos - b, side="left")
        hi = np.searchsorted(cnd_pos, v_pos - a, side="left")
        sel = cnd_pos[lo:hi]
        left.append(sel[-k:] if mode == "near" else sel[:k])
        # right bin: v+a < p <= v+b ; 'near' = first k, 'far' = last k
        lo = np.searchsorted(cnd_pos, v_pos + a, side="right")
        hi = np.searchsorted(cnd_pos, v_pos + b, side="right")
        sel = cnd_pos[lo:hi]
        right.append(sel[:k] if mode == "near" else sel[-k:])
    ls1 = sorted(int(x) for x in np.concatenate(left)) if left else []
    ls2 = sorted(int(x) for x in np.concatenate(right)) if right else []
    return ls1, ls2


def expand_reads(rs, lo, hi, flag_filter, batch=512):
    """Pileup entries (p0, read_index, code) of all admitted reads over 0-based [lo, hi), in
    (read order) — restates the column contract of SURVEY.md Appendix C.4-5 for token[0] only:
    M/=/X -> base code of the query base, D/N -> 4."""
    sel = np.nonzero((rs.pos < hi) & (rs.ref_end > lo) & ((rs.flag & flag_filter) == 0) & ((rs.flag & 4) == 0))[0]
    Ps, Rs, Cs = [], [], []
    for b0 in range(0, len(sel), batch):
        idx = sel[b0:b0 + batch]
        nc = (rs.cigar_off[idx + 1] - rs.cigar_off[idx]).astype(np.int64)
        tot = int(nc.sum())
        if tot == 0:
            continue
        rid = np.repeat(np.arange(len(idx)), nc)
        first = np.cumsum(nc) - nc
        cpos = np.arange(tot) - np.repeat(first, nc) + np.repeat(rs.cigar_off[idx], nc)
        cg = rs.cigar[cpos]
        ops = (cg & 15).astype(np.int64)
        lens = (cg >> 4).astype(np.int64)
        rl = lens * _REF_CONSUME[ops]
        ql = lens * _QRY_CONSUME[ops]
        crl = np.cumsum(rl) - rl
        cql = np.cumsum(ql) - ql
        rstart = rs.pos[idx][rid].astype(np.int64) + crl - crl[first][rid]
        qstart = cql - cql[first][rid]
        nz = np.nonzero(rl > 0)[0]
        opidx = np.repeat(nz, rl[nz])
        off = np.arange(len(opidx)) - np.repeat(np.cumsum(rl[nz]) - rl[nz], rl[nz])
        p = rstart[opidx] + off
        keep = (p >= lo) & (p < hi)
        opidx, off, p = opidx[keep], off[keep], p[keep]
        o = ops[opidx]
        is_m = (o == 0) | (o == 7) | (o == 8)
        r_local = rid[opidx]
        ridx = idx[r_local]
        q = qstart[opidx] + off
        code = np.full(len(p), 4, np.int8)
        qm = q[is_m]
        rm = ridx[is_m]
        inb = qm < rs.l_seq[rm]
        byte = rs.seq4[np.minimum(rs.seq_off[rm] + (qm >> 1), len(rs.seq4) - 1)]
        nib = np.where((qm & 1) == 0, byte >> 4, byte & 15)
        code[is_m] = np.where(inb, NIB_TO_CODE[nib], 4)
        Ps.append(p); Rs.append(ridx); Cs.append(code)
    if not Ps:
        z = np.zeros(0, np.int64)
        return z, z, np.zeros(0, np.int8)
    return np.concatenate(Ps), np.concatenate(Rs), np.concatenate(Cs)


def get_snp_testing_candidates(rs, dct, region, bed_intervals=None, return_aux=False):
    """generate_SNP_pileups.py:103-278.  `rs` replaces sam_path/fasta_path; `bed_intervals`
    (list of (start, end) for this contig, or None) replaces the tabix exclude file."""
    start, end, ploidy = region["start"], region["end"], region["ploidy"]
    threshold = dct["threshold"]
    nbr_size = 20
    L = rs.contig_len

    # :137  ref_dict over v_pos in [max(1,start-50000), end+50000] (clipped by the contig)
    v_lo = max(1, start - 50000)
    ref_code_at = lambda v: int(REF_TO_CODE[rs.ref[v - 1]])

    # :139-143  strand of every non-secondary, non-supplementary read overlapping [start-10, end+10)
    s_lo, s_hi = max(0, start - 10), end + 10
    strand = ((rs.flag & 0x10) >> 4).astype(np.int8)

    flag = 0x4 | 0x100 | 0x200 | 0x400 if dct.get("supplementary") else 0x4 | 0x100 | 0x200 | 0x400 | 0x800

    # :156  pileup window, 0-based half-open, truncate=True
    lo, hi = max(0, start - 1 - 50000), min(end + 50000, L)
    W = max(0, hi - lo)
    P, R, C = expand_reads(rs, lo, hi, flag)
    rel = (P - lo).astype(np.int64)
    cnt = np.bincount(rel * 5 + C, minlength=W * 5).reshape(W, 5) if W else np.zeros((0, 5), np.int64)
    n = cnt.sum(1)                                   # :164 get_num_aligned (deletions and N included)
    refc = REF_TO_CODE[rs.ref[lo:hi]].astype(np.int64)  # :159
    ok = (refc < 4) & (n > 0)
    if bed_intervals:                                # :113-126,161 — 1-based v_pos against [bed_start, bed_end)
        vp = np.arange(lo + 1, hi + 1)
        ex = np.zeros(W, bool)
        for bs, be in bed_intervals:
            ex |= (vp >= bs) & (vp < be)
        ok &= ~ex
    acgt = cnt[:, :4].copy()
    acgt[np.arange(W), np.minimum(refc, 3)] = np.where(refc < 4, 0, acgt[np.arange(W), np.minimum(refc, 3)])
    alt = acgt.max(1) if W else np.zeros(0, np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        alt_freq = alt / n                           # :166 python int / int -> correctly rounded float64
    cov_ok = ok & (n >= dct["mincov"])               # :170
    if ploidy == "diploid":                          # :172-175
        is_nbr = cov_ok & (threshold[0] <= alt_freq) & (alt_freq < threshold[1])
    else:                                            # :176-179
        is_nbr = cov_ok & (threshold[0] <= alt_freq)
    vpos_all = np.arange(lo + 1, hi + 1)
    is_cand = cov_ok & (vpos_all >= start) & (vpos_all <= end) & (dct["min_allele_freq"] <= alt_freq)  # :183
    nbr_sites = vpos_all[is_nbr]
    cand_sites = vpos_all[is_cand]

    # per kept site: {read -> code}; duplicate qnames collapse, last one wins (:175,:185)
    kept = is_nbr | is_cand
    m = kept[rel]
    kp, kr, kc = P[m] + 1, R[m], C[m]
    order = np.argsort(kp, kind="stable")
    kp, kr, kc = kp[order], kr[order], kc[order]
    ukp, first = np.unique(kp, return_index=True)
    bounds = np.concatenate([first, [len(kp)]])
    col_of = {int(v): (int(bounds[i]), int(bounds[i + 1])) for i, v in enumerate(ukp)}
    dup_names = rs.qnames is not None and len(set(rs.qnames)) != len(rs.qnames)

    def column(v):
        a, b = col_of[v]
        reads, codes = kr[a:b], kc[a:b]
        if dup_names:
            d = {}
            for r_, c_ in zip(reads, codes):
                d[rs.qname(int(r_))] = (int(r_), int(c_))
            reads = np.array([x[0] for x in d.values()], np.int64)
            codes = np.array([x[1] for x in d.values()], np.int8)
        return reads, codes

    out_pos, out_ref, out_mat, out_dp, out_freq, out_fwd, out_rev, cur_depth = [], [], [], [], [], [], [], []
    aux_cols = []
    for v_pos in cand_sites:
        v_pos = int(v_pos)
        ls1, ls2 = get_cnd_pos(v_pos, nbr_sites, dct["seq"])           # :202
        reads, codes = column(v_pos)
        onehot = np.eye(5)[codes][:, :4]                               # :210
        st = strand[reads].astype(bool)                                # :211
        fwd_bases = onehot[~st].sum(0)                                 # :212
        rev_bases = onehot[st].sum(0)                                  # :213
        if len(reads) > dct["maxcov"]:
            # :215-216 is random.sample with an UNSEEDED generator in the reference.
            # Deterministic rule shared with the CUDA path: keep the maxcov earliest reads in BAM order.
            keep_idx = np.argsort(reads, kind="stable")[:dct["maxcov"]]
            keep_idx.sort()
            reads, codes = reads[keep_idx], codes[keep_idx]
        cols = ls1 + [v_pos] + ls2
        tmp = np.full((len(reads), len(cols)), 4, np.int64)            # :221
        pos_in_sample = {int(r_): i for i, r_ in enumerate(reads)}
        for j, nb in enumerate(cols):                                  # :223-237
            if nb == v_pos:
                tmp[:, j] = codes
                continue
            nr, ncodes = column(nb)
            for r_, c_ in zip(nr, ncodes):
                i = pos_in_sample.get(int(r_))
                if i is not None:
                    tmp[i, j] = c_
        total_rlist = np.array([ref_code_at(c) for c in cols])         # :238-242
        if len(total_rlist) < dct["min_nbr_sites"]:                    # :244
            continue
        cc = len(ls1)
        mat = np.stack([np.eye(5)[tmp[tmp[:, cc] == i]].sum(0) for i in range(4)])[:, :, :4]   # :247
        total_ref = np.eye(5)[total_rlist]                             # :249
        total_ref[:, 4] = 0                                            # :250
        total_ref = total_ref[np.newaxis, :]
        mat = np.dstack([mat, np.zeros([4, mat.shape[1]]) + np.eye(4)[ref_code_at(v_pos)][:, np.newaxis]])  # :252
        data = np.vstack([total_ref, mat * (1 - 2 * total_ref)])      # :253
        data = np.hstack([np.zeros([5, nbr_size - len(ls1), 5]), data,
                          np.zeros([5, nbr_size - len(ls2), 5])]).astype(np.int32)             # :254
        out_pos.append(v_pos); out_ref.append(ref_code_at(v_pos)); out_mat.append(data)
        a, b = col_of[v_pos]
        out_dp.append(int(n[v_pos - 1 - lo])); out_freq.append(float(alt_freq[v_pos - 1 - lo]))
        out_fwd.append(fwd_bases); out_rev.append(rev_bases); cur_depth.append(len(reads))
        if return_aux:
            aux_cols.append(cols)
    depth = 0
    if len(out_pos) > 0:                                               # :265-277
        out_mat = np.array(out_mat).astype(np.float32)
        out_pos = np.array(out_pos)
        out_ref = np.eye(max(4, np.max(out_ref) + 1))[np.array(out_ref)].astype(np.int32)[:, :4]
        out_dp = np.array(out_dp)
        out_freq = np.array(out_freq)
        depth = np.mean(cur_depth)
        out_fwd = np.array(out_fwd)
        out_rev = np.array(out_rev)
    res = (out_pos, out_ref, out_mat, out_dp, out_freq, depth, out_fwd, out_rev)
    if return_aux:
        return res, {"nbr_sites": nbr_sites, "cand_sites": cand_sites, "cols": aux_cols,
                     "sample_depth": np.array(cur_depth, np.int64)}
    return res


def scale_counts(x_test, train_coverage, coverage=None, dp=None):
    """N1 — snpCaller.py:90-96 (diploid) / :167-173 (haploid), with the float semantics of the
    reference's pinned environment (environment.yml:9 numpy<2): with a scalar `coverage` the float64
    ratio is demoted and the multiply is done once in float32; with `dp` (--disable_coverage_
    normalization) the ratio is a float64 ARRAY, so the product is float64 and then rounded to float32."""
    x = np.array(x_test, dtype=np.float32, copy=True)
    if dp is not None:
        ratio = float(train_coverage) / np.asarray(dp)[:, np.newaxis, np.newaxis, np.newaxis].astype(np.float64)
        x[:, 1:, :, :4] = (x[:, 1:, :, :4].astype(np.float64) * ratio).astype(np.float32)
    else:
        s = np.float32(float(train_coverage) / float(coverage))
        x[:, 1:, :, :4] = x[:, 1:, :, :4] * s
    return x
