"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The two third-party algorithms under the reference's indel path are external binaries/libraries that are neither
vendored nor installable here: MUSCLE 3.8 (`generate_indel_pileups.py:30`) and parasail
(`generate_indel_pileups.py:79`).  Their numerics cannot be pinned, so this repo DEFINES its own deterministic
replacements, states them here in plain numpy, and holds the CUDA kernels bit-exact to THESE definitions:

  star_msa      reference-anchored star alignment, stands in for `muscle -maxiters 1 -diags1 -gapopen 1.0`
  nw_trace      global affine-gap alignment with traceback, stands in for `parasail.nw_trace(q, r, 9, 1, M(20,-10))`

`oracle/shim/bin/muscle` and `oracle/shim/parasail.py` wrap them so the UNMODIFIED reference modules run on top.
PARITY UNPINNED against MUSCLE / parasail (by construction).  Integer arithmetic only; every tie is broken by a
stated rule.

star_msa(reads, ref)
  1. every read slice is aligned to the reference slice by global linear-gap Needleman-Wunsch:
       match +2, mismatch -4, gap -3 per base; H[0][0] = 0, first row / column = -3 k;
       H[i][j] = max(H[i-1][j-1] + s(read[i-1], ref[j-1]),  H[i-1][j] - 3,  H[i][j-1] - 3)
     direction of a cell: DIAG if the diagonal candidate attains the maximum, else UP (read base unaligned = insertion)
     if the up candidate attains it, else LEFT (reference base unaligned = deletion); traceback from (n, m).
  2. columns: one per reference base, plus after reference base j (j = -1: before the first) as many insertion
     columns as the longest insertion any read of the group has there; shorter insertions are left-aligned in the
     block and padded with '-'; deleted reference bases are '-' in the read's row; the reference row has '-' in
     insertion columns.
"""
import numpy as np

MATCH, MISMATCH, GAP = 2, -4, 3
_CODE = {"A": 0, "G": 1, "T": 2, "C": 3}
DIAG, UP, LEFT = 0, 1, 2


def encode(s):
    return np.array([_CODE.get(c, 4) for c in s], dtype=np.int64)


def nw_linear(read, ref):
    """-> (dir matrix uint8 [n+1][m+1], score).  read/ref: int code arrays (4 = N matches nothing)."""
    n, m = len(read), len(ref)
    H_prev = -GAP * np.arange(m + 1, dtype=np.int64)
    D = np.zeros((n + 1, m + 1), np.uint8)
    D[0, 1:] = LEFT
    jj = np.arange(m + 1, dtype=np.int64)
    for i in range(1, n + 1):
        s = np.where((ref == read[i - 1]) & (read[i - 1] < 4), MATCH, MISMATCH)
        diag = H_prev[:-1] + s                       # j = 1..m
        up = H_prev[1:] - GAP
        c = np.maximum(diag, up)
        cand = np.concatenate([[-GAP * i], c])       # cand[0] = H[i][0]
        # H[i][j] = max_k<=j (cand[k] - GAP (j-k))
        H = np.maximum.accumulate(cand + GAP * jj) - GAP * jj
        d = np.where(H[1:] == diag, DIAG, np.where(H[1:] == up, UP, LEFT)).astype(np.uint8)
        D[i, 1:] = d
        D[i, 0] = UP
        H_prev = H
    return D, int(H_prev[m])


def traceback_linear(D, n, m):
    """-> (aligned[m] code of the read at each reference base: 0..3, 4 = N, 5 = deleted,
           ins[m+1]: list of read base codes inserted after reference base j-1 (index 0 = before the first))."""
    aligned = np.full(m, 5, np.int64)
    ins = [[] for _ in range(m + 1)]
    i, j = n, m
    steps = []
    while i > 0 or j > 0:
        d = D[i, j]
        steps.append(d)
        if d == DIAG:
            i -= 1; j -= 1
        elif d == UP:
            i -= 1
        else:
            j -= 1
    return steps[::-1]


def align_read(read, ref):
    """read, ref: code arrays.  -> (aligned[m], ins list of lists[m+1])"""
    n, m = len(read), len(ref)
    D, _ = nw_linear(read, ref)
    steps = traceback_linear(D, n, m)
    aligned = np.full(m, 5, np.int64)
    ins = [[] for _ in range(m + 1)]
    i = j = 0
    for d in steps:
        if d == DIAG:
            aligned[j] = read[i]; i += 1; j += 1
        elif d == UP:
            ins[j].append(int(read[i])); i += 1
        else:
            j += 1
    return aligned, ins


def star_columns(alignments, m):
    """alignments: list of (aligned, ins).  -> (rows int [n_reads][L] with 0..3 base, 4 gap (N folded to gap is NOT done: N = 6),
    ref_cols: index of the MSA column of every reference base, L)"""
    width = np.zeros(m + 1, np.int64)
    for _, ins in alignments:
        for j in range(m + 1):
            if len(ins[j]) > width[j]:
                width[j] = len(ins[j])
    # column of reference base j = j + sum(width[0..j])
    off = np.cumsum(width)
    L = m + int(off[m])
    ref_col = np.arange(m) + off[:m]
    rows = np.full((len(alignments), L), 4, np.int64)
    for r, (aligned, ins) in enumerate(alignments):
        for j in range(m + 1):
            start = (ref_col[j - 1] + 1) if j > 0 else 0
            for k, b in enumerate(ins[j]):
                rows[r, start + k] = b
        a = aligned.copy()
        a[a == 5] = 4
        rows[r, ref_col] = a
    return rows, ref_col, L


def star_msa(reads, ref):
    """reads: list of strings, ref: string -> (aligned read strings, aligned ref string)."""
    sym = np.array(list("AGTC-"))
    rc = encode(ref)
    als = [align_read(encode(r), rc) for r in reads]
    rows, ref_col, L = star_columns(als, len(rc))
    ref_row = np.full(L, 4, np.int64)
    ref_row[ref_col] = rc
    return ["".join(sym[np.minimum(x, 4)]) for x in rows], "".join(sym[np.minimum(ref_row, 4)])


# ------------------------------------------------------------------------------------------------
# nw_trace: global alignment, affine gaps (a gap of length L costs open + (L-1) * extend), match/mismatch matrix.
#   H = best score ending at (i, j); E = ending with a gap in the query (reference base unaligned, CIGAR 'D');
#   F = ending with a gap in the reference (query base unaligned, CIGAR 'I').
#   H[i][j] = max(H[i-1][j-1] + s, F[i][j], E[i][j]); ties: DIAG, then F ('I'), then E ('D').
#   E[i][j] = max(E[i][j-1] - extend, H[i][j-1] - open); ties: extend.   F likewise along i.
# Traceback from (n, m) in state H.
# ------------------------------------------------------------------------------------------------
NEG = -(1 << 40)


def nw_trace(query, ref, gap_open=9, gap_extend=1, match=20, mismatch=-10):
    """query, ref: strings over AGTC.  -> list of (op, length) with op in '=XID' (query vs ref)."""
    q, r = encode(query), encode(ref)
    n, m = len(q), len(r)
    H = np.full((n + 1, m + 1), NEG, np.int64)
    E = np.full((n + 1, m + 1), NEG, np.int64)
    F = np.full((n + 1, m + 1), NEG, np.int64)
    H[0, 0] = 0
    for j in range(1, m + 1):
        E[0, j] = -gap_open - gap_extend * (j - 1); H[0, j] = E[0, j]
    for i in range(1, n + 1):
        F[i, 0] = -gap_open - gap_extend * (i - 1); H[i, 0] = F[i, 0]
    jj = np.arange(m + 1, dtype=np.int64)
    for i in range(1, n + 1):
        s = np.where(r == q[i - 1], match, mismatch)
        diag = H[i - 1, :-1] + s
        F[i, 1:] = np.maximum(F[i - 1, 1:] - gap_extend, H[i - 1, 1:] - gap_open)
        T = np.maximum(diag, F[i, 1:])                                    # candidates not coming from E, j = 1..m
        Tfull = np.concatenate([[H[i, 0]], T])
        # E[i][j] = max_{k<j} (Tfull[k] - open - extend (j-k-1))   (opening from an E-derived H is dominated since open >= extend)
        scan = np.maximum.accumulate(Tfull + gap_extend * jj)
        E[i, 1:] = scan[:-1] - gap_open - gap_extend * (jj[1:] - 1)
        H[i, 1:] = np.maximum(T, E[i, 1:])
    # traceback
    ops = []
    i, j, state = n, m, "H"
    while i > 0 or j > 0:
        if state == "H":
            if i > 0 and j > 0 and H[i, j] == H[i - 1, j - 1] + (match if q[i - 1] == r[j - 1] else mismatch):
                ops.append("=" if q[i - 1] == r[j - 1] else "X"); i -= 1; j -= 1
            elif i > 0 and H[i, j] == F[i, j]:
                state = "F"
            else:
                state = "E"
        elif state == "F":
            ops.append("I")
            if i > 1 and F[i, j] == F[i - 1, j] - gap_extend:
                i -= 1
            else:
                i -= 1; state = "H"
        else:
            ops.append("D")
            if j > 1 and E[i, j] == E[i, j - 1] - gap_extend:
                j -= 1
            else:
                j -= 1; state = "H"
    ops = ops[::-1]
    out = []
    for o in ops:
        if out and out[-1][0] == o:
            out[-1][1] += 1
        else:
            out.append([o, 1])
    return [(o, l) for o, l in out]


CIGAR_CODE = {"=": 7, "X": 8, "I": 1, "D": 2}


def cigar_words(ops):
    return [(l << 4) | CIGAR_CODE[o] for o, l in ops]
