"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's diploid indel feature path over an in-memory `ReadSet`
(paths under /root/reference/nanocaller_src/):

  get_indel_testing_candidates   generate_indel_pileups.py:129-370   (incl. the impute_indel_phase branch :278-304, :309-313)
  msa (tensor / consensus part)  generate_indel_pileups.py:12-73     on top of oracle/star_msa.star_msa  (stands in for MUSCLE)
  allele_prediction              generate_indel_pileups.py:77-127    on top of oracle/star_msa.nw_trace  (stands in for parasail)

Pinned: tests/test_indel_oracle_golden.py compares every output with fixtures produced by the UNMODIFIED reference
module run over oracle/shim (tests/golden/make_golden_indel.py).  MUSCLE / parasail themselves are unpinned
(oracle/star_msa.py explains the stand-ins); htslib column semantics as in oracle/shim/pysam.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module.
"""
import numpy as np

from . import star_msa

_REF_CONSUME = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1] + [0] * 7, dtype=np.int64)
_QRY_CONSUME = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1] + [0] * 7, dtype=np.int64)
_NIB = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)


def read_events(rs, i):
    """Indel annotations of read i (appendix C.4): list of (p0 of the column carrying the token, +L / -L)."""
    cg = rs.read_cigar(i)
    ops = (cg & 15).astype(np.int64)
    lens = (cg >> 4).astype(np.int64)
    rl = lens * _REF_CONSUME[ops]
    rstart = int(rs.pos[i]) + np.concatenate([[0], np.cumsum(rl)[:-1]])
    nc = len(ops)
    ev = []
    for k in range(nc):
        if rl[k] == 0 or k + 1 >= nc:
            continue
        op, op2 = ops[k], ops[k + 1]
        tot = 0
        if op2 == 2 and op != 2:
            tot = lens[k + 1]
            j = k + 2
            while j < nc:
                if ops[j] == 2:
                    tot += lens[j]
                elif ops[j] in (1, 4, 0, 7, 8):
                    break
                j += 1
            tot = -tot
        elif op2 == 1 or (op2 == 6 and k + 2 < nc):
            j = k + 1
            while j < nc:
                if ops[j] == 1:
                    tot += lens[j]
                elif ops[j] != 6:
                    break
                j += 1
        if tot:
            ev.append((int(rstart[k] + rl[k] - 1), int(tot)))
    return ev


def qpos_or_next(rs, i, p0):
    """query_position_or_next of read i at column p0 (must be covered)."""
    cg = rs.read_cigar(i)
    x, y = int(rs.pos[i]), 0
    for w in cg:
        op, ln = int(w & 15), int(w >> 4)
        r, q = ln * _REF_CONSUME[op], ln * _QRY_CONSUME[op]
        if r and x <= p0 < x + r:
            return y + (p0 - x) if q else y
        x += r
        y += q
    raise ValueError("position not covered")


def msa_tensor(seqs, ref, mincov, maxcov):
    """generate_indel_pileups.py:12-73 with the star alignment in place of MUSCLE.
    -> (flag, tensor float64 [5,128,2] | None, consensus | None, ref_seq | None)"""
    sample = list(seqs)
    if len(sample) > maxcov:
        sample = sample[:maxcov]                                  # :19 is an UNSEEDED random.sample: first maxcov in pileup order
    if len(sample) < mincov:                                      # :48
        return 0, None, None, None
    rows, ref_row = star_msa.star_msa(sample, ref)
    mapping = {"A": 0, "G": 1, "T": 2, "C": 3, "-": 4}
    ref_mat = np.eye(5)[[mapping[x] for x in ref_row]]            # :54
    mat = np.array([[mapping[c] for c in x] for x in rows])
    h0 = np.sum(np.eye(5)[mat], axis=0).astype(np.float32)        # :58
    alt = h0 / np.sum(h0, axis=1)[:, np.newaxis]                  # :59 float32
    tmp = np.copy(alt)
    tmp[:, 4] = tmp[:, 4] - np.float32(0.01)                      # :62
    cns = "".join("AGTC-"[x] for x in np.argmax(tmp, axis=1)).replace("-", "")
    ref_seq = ref_row.replace("-", "")
    alt -= ref_mat.astype(np.float32)                             # :67 (in place on a float32 array)
    final = np.dstack([alt, ref_mat])[:128, :, :].transpose(1, 0, 2)
    if final.shape[1] < 128:
        final = np.hstack((final, np.zeros((5, 128 - final.shape[1], 2))))
    return 1, final, cns, ref_seq


def allele_prediction(alt, ref_seq, max_range):
    """generate_indel_pileups.py:77-127 (control flow kept line for line)."""
    cigar_op = [(star_msa.CIGAR_CODE[o], l) for o, l in star_msa.nw_trace(alt, ref_seq, 9, 1, 20, -10)]
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


def read_token(rs, i, p0):
    """Upper-cased pileup string of read i at column p0 with add_indels=True (appendix C.4; generate_indel_pileups.py:279):
    base letter / '*' / '<' '>' and, on the last column of a CIGAR op that is followed by an insertion or deletion,
    '+L<inserted bases>' or '-L' + 'N' * L."""
    cg = rs.read_cigar(i)
    x, y = int(rs.pos[i]), 0
    for k, w in enumerate(cg):
        op, ln = int(w & 15), int(w >> 4)
        r, q = ln * int(_REF_CONSUME[op]), ln * int(_QRY_CONSUME[op])
        if r and x <= p0 < x + r:
            nib = rs.read_nibbles(i)
            if op in (0, 7, 8):
                qi = y + (p0 - x)
                tok = chr(_NIB[nib[qi]]) if qi < len(nib) else "N"
            elif op == 2:
                tok = "*"
            else:
                tok = "<" if (int(rs.flag[i]) & 0x10) else ">"
            if p0 == x + r - 1:
                L = dict(read_events(rs, i)).get(p0, 0)
                if L > 0:
                    qn = y + q                                   # P ops between the insertions consume no query
                    tok += "+%d%s" % (L, _NIB[nib[qn:qn + L]].tobytes().decode())
                elif L < 0:
                    tok += "-%d%s" % (-L, "N" * (-L))
            return tok
        x += r
        y += q
    raise ValueError("position not covered")


def impute_column(tokens, mincov):
    """generate_indel_pileups.py:287-300 for one column: `tokens` = upper-cased pileup strings in pileup order.
    -> None, or (indices of read_names_0, indices of read_names_1)."""
    n = len(tokens)
    groups = {}
    for k, s in enumerate(tokens):
        groups.setdefault(s, []).append(k)
    counts = sorted([(x, len(groups[x])) for x in groups], key=lambda x: x[1], reverse=True)
    if counts[0][1] <= 0.8 * n:
        names0 = list(groups[counts[0][0]])
        names1 = list(groups[counts[1][0]]) if counts[1][1] >= mincov else [k for k in range(n) if k not in set(names0)]
    else:
        names0 = groups[counts[0][0]][:counts[0][1] // 2]
        names1 = groups[counts[0][0]][counts[0][1] // 2:]
    if len(names0) >= mincov and len(names1) >= mincov:
        return names0, names1
    return None


def scan_variants(rs, dct, chunk, bed_intervals=None, haploid=False, extra_out=None):
    """Pass 1 (generate_indel_pileups.py:213-304; haploid: generate_indel_pileups_haploid.py:199-241):
    -> dict {key v_pos: type 0 | 1}.  With dct['impute_indel_phase'], `extra_out` (a dict) receives
    extra_variants {key: (read indices of read_names_0, of read_names_1)} (:278-304)."""
    start, end = chunk["start"], chunk["end"]
    W, SW = dct["win_size"], dct["small_win_size"]
    mincov, ins_t, del_t = dct["mincov"], dct["ins_t"], dct["del_t"]
    flag = (0x4 | 0x100 | 0x200 | 0x400) if dct.get("supplementary") else (0x4 | 0x100 | 0x200 | 0x400 | 0x800)
    lo, hi = max(0, start - 1), min(end, rs.contig_len)
    n = hi - lo
    if n <= 0:
        return {}
    adm = np.nonzero(((rs.flag & flag) == 0) & (rs.pos < hi) & (rs.ref_end > lo) & (rs.ref_end > rs.pos))[0]
    depth = np.zeros((3, n + 1), np.int64)           # hap0, hap1, all
    ev_pos = {k: [] for k in range(8)}               # (hap, kind) -> list of (rank-space handled later) (p0, read)
    impute = bool(dct.get("impute_indel_phase")) and not haploid
    col_del = np.zeros(n + 1, np.int64)              # '*' and '-' characters among the first two of every read's string (:283)
    col_ins = np.zeros(n + 1, np.int64)              # '+' characters
    for i in adm:
        a, b = max(lo, int(rs.pos[i])) - lo, min(hi, int(rs.ref_end[i])) - lo
        depth[2, a] += 1; depth[2, b] -= 1
        if impute:
            cg = rs.read_cigar(int(i))
            x = int(rs.pos[i])
            for w in cg:
                op, ln = int(w & 15), int(w >> 4)
                if op == 2:
                    col_del[max(lo, x) - lo:max(0, min(hi, x + ln) - lo)] += 1
                x += ln * int(_REF_CONSUME[op])
            for p0, L in read_events(rs, int(i)):
                if lo <= p0 < hi:
                    (col_del if L < 0 else col_ins)[p0 - lo] += 1
        h = 0 if haploid else int(rs.hp[i]) - 1
        if h in (0, 1):
            depth[h, a] += 1; depth[h, b] -= 1
            for p0, L in read_events(rs, int(i)):
                if not (lo <= p0 < hi):
                    continue
                kind_big = 2 < abs(L) <= 50
                kind_small = abs(L) <= 10
                base = 0 if L < 0 else 2             # del: 0 (large) 1 (small); ins: 2 (large) 3 (small)
                if kind_big:
                    ev_pos[h * 4 + base].append((p0, int(i)))
                if kind_small:
                    ev_pos[h * 4 + base + 1].append((p0, int(i)))
    depth = np.cumsum(depth[:, :n], axis=1)
    emitted = depth[2] > 0
    if bed_intervals:
        vp = np.arange(lo + 1, hi + 1)
        for bs, be in bed_intervals:
            emitted &= ~((vp >= bs) & (vp < be))
    rank = np.cumsum(emitted) - 1                    # rank of column among emitted, valid where emitted
    n_em = int(emitted.sum())
    union = np.zeros((8, n_em + 1), np.int64)
    for key, lst in ev_pos.items():
        win = W if key % 2 == 0 else SW
        per_read = {}
        for p0, i in lst:
            if emitted[p0 - lo]:
                per_read.setdefault(i, []).append(int(rank[p0 - lo]))
        for i, rks in per_read.items():
            rks.sort()
            cur_a, cur_b = None, None
            for r in rks:                            # union of [r, r+win-1] intervals of this read
                a, b = r, min(n_em, r + win)
                if cur_a is None:
                    cur_a, cur_b = a, b
                elif a <= cur_b:
                    cur_b = max(cur_b, b)
                else:
                    union[key, cur_a] += 1; union[key, cur_b] -= 1
                    cur_a, cur_b = a, b
            if cur_a is not None:
                union[key, cur_a] += 1; union[key, cur_b] -= 1
    union = np.cumsum(union[:, :n_em], axis=1)
    variants = {}
    prev = 0
    cols = np.nonzero(emitted)[0]
    for r, c in enumerate(cols):
        v_pos = lo + int(c) + 1
        if v_pos <= prev:
            continue
        l0, l1 = int(depth[0, c]), int(depth[1, c])
        if haploid:
            if l0 >= mincov:
                f = lambda key: union[key, r] / l0 if l0 > 0 else 0
                if f(0) >= del_t or f(2) >= ins_t:
                    prev = v_pos + W
                    variants[max(1, v_pos - W)] = 0
                elif f(1) >= del_t or f(3) >= ins_t or (f(1) + f(3)) >= 0.9:
                    prev = v_pos + 10
                    variants[max(1, v_pos - 10)] = 1
        elif l0 >= mincov and l1 >= mincov:
            f = lambda key, l: union[key, r] / l if l > 0 else 0
            del0, dels0, ins0, inss0 = f(0, l0), f(1, l0), f(2, l0), f(3, l0)
            del1, dels1, ins1, inss1 = f(4, l1), f(5, l1), f(6, l1), f(7, l1)
            if max([del0, del1]) >= del_t or max([ins0, ins1]) >= ins_t:
                prev = v_pos + W
                variants[max(1, v_pos - W)] = 0
            elif max([dels0, dels1]) >= del_t or max([inss0, inss1]) >= ins_t or (dels0 + inss0) >= 0.9 or (dels1 + inss1) >= 0.9:
                prev = v_pos + 10
                variants[max(1, v_pos - 10)] = 1
        elif impute and int(depth[2, c]) >= 2 * mincov:                      # :278
            ltot = int(depth[2, c])
            if del_t <= col_del[c] / ltot or ins_t <= col_ins[c] / ltot:     # :283-285
                p0 = lo + int(c)
                cov = [int(i) for i in adm if rs.pos[i] <= p0 < rs.ref_end[i]]
                assert len(cov) == ltot
                got = impute_column([read_token(rs, i, p0) for i in cov], mincov)
                if got is not None:
                    prev = v_pos + 10
                    variants[max(1, v_pos - 10)] = 1
                    if extra_out is not None:
                        extra_out[max(1, v_pos - 10)] = ([cov[k] for k in got[0]], [cov[k] for k in got[1]])
    return variants


def site_slices(rs, dct, chunk, v_pos):
    """Pass 2 for one key position (generate_indel_pileups.py:306-338): -> None if the column is not emitted or the
    reference window has a non-ACGT base, else (ref string, [(read index, hp, slice string)] in pileup order)."""
    window_after = 260 if dct["seq"] == "pacbio" else 160
    start, end = chunk["start"], chunk["end"]
    flag = (0x4 | 0x100 | 0x200 | 0x400) if dct.get("supplementary") else (0x4 | 0x100 | 0x200 | 0x400 | 0x800)
    p0 = v_pos - 1
    if not (max(0, start - 10 - dct["win_size"]) <= p0 < min(end, rs.contig_len)):
        return None
    cov = np.nonzero(((rs.flag & flag) == 0) & (rs.pos <= p0) & (rs.ref_end > p0))[0]
    if len(cov) == 0:
        return None
    L = rs.contig_len
    k_lo, k_hi = max(1, start - 200), min(end + 400, L)                     # ref_dict keys (:174)
    ref_chars = []
    for p in range(v_pos, min(L, v_pos + window_after + 1)):
        if not (k_lo <= p <= k_hi):
            raise KeyError(p)
        c = chr(rs.ref[p - 1])
        ref_chars.append(c if c in "AGTC" else "N")
    ref = "".join(ref_chars)
    if "N" in ref:
        return None
    out = []
    for i in cov:
        q = qpos_or_next(rs, int(i), p0)
        nib = rs.read_nibbles(int(i))
        out.append((int(i), int(rs.hp[i]), _NIB[nib[q:q + window_after]].tobytes().decode()))
    return ref, out


def get_indel_testing_candidates(rs, dct, chunk, bed_intervals=None):
    """generate_indel_pileups.py:129-370 -> (pos, x0, x1, x2, alleles, phase)."""
    W = dct["win_size"]
    max_range = {0: max(10, W), 1: 10}
    extra = {}
    variants = scan_variants(rs, dct, chunk, bed_intervals, extra_out=extra)
    pos, X0, X1, X2, alleles, phase = [], [], [], [], [], []
    for v_pos in sorted(variants):
        ss = site_slices(rs, dct, chunk, v_pos)
        if ss is None:
            continue
        ref, reads = ss
        tot = [s for _, _, s in reads]
        if v_pos in extra:                                                   # :309-313: the imputed read sets win over the HP tags
            n0, n1 = set(extra[v_pos][0]), set(extra[v_pos][1])
            h0 = [(i, s) for i, hp, s in reads if i in n0]
            h1 = [(i, s) for i, hp, s in reads if i not in n0 and i in n1]
        else:
            h0 = [(i, s) for i, hp, s in reads if hp == 1]
            h1 = [(i, s) for i, hp, s in reads if hp == 2]
        f0, d0, a0, r0 = msa_tensor([s for _, s in h0], ref, 2, dct["maxcov"])
        f1, d1, a1, r1 = msa_tensor([s for _, s in h1], ref, 2, dct["maxcov"])
        ft, dt, at, rt = msa_tensor(tot, ref, dct["mincov"], dct["maxcov"])
        if f0 and f1 and ft:
            pos.append(v_pos); X0.append(d0); X1.append(d1); X2.append(dt)
            phase.append(int(rs.ps[h0[0][0]]) if rs.hp[h0[0][0]] > 0 else None)     # phase_dict (:180-186): None without an HP tag
            mr = max_range[variants[v_pos]]
            alleles.append([allele_prediction(a0, r0, mr), allele_prediction(a1, r1, mr), allele_prediction(at, rt, mr)])
    if not pos:
        return pos, X0, X1, X2, alleles, phase
    return pos, np.array(X0), np.array(X1), np.array(X2), alleles, phase


def get_indel_testing_candidates_haploid(rs, dct, chunk, bed_intervals=None):
    """generate_indel_pileups_haploid.py:129-277 -> (pos, x_total, alleles)."""
    W = dct["win_size"]
    max_range = {0: max(10, W), 1: 10}
    variants = scan_variants(rs, dct, chunk, bed_intervals, haploid=True)
    pos, X, alleles = [], [], []
    for v_pos in sorted(variants):
        ss = site_slices(rs, dct, chunk, v_pos)
        if ss is None:
            continue
        ref, reads = ss
        ft, dt, at, rt = msa_tensor([s for _, _, s in reads], ref, dct["mincov"], dct["maxcov"])
        if ft:
            pos.append(v_pos); X.append(dt)
            alleles.append(allele_prediction(at, rt, max_range[variants[v_pos]]))
    if not pos:
        return pos, X, alleles
    return pos, np.array(X), alleles
