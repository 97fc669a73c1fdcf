"""GPU: the aligned-row fill (K0b, `row_fill_kernel`) on reads built to hit its special paths — every start offset within a row
word, all nine CIGAR operations, blocks with more operations than the warp stages (per-word search through global memory), a read
longer than the per-read block index (every word searches all operations), a read whose sequence is shorter than its CIGAR says —
checked through the public pileup call against the oracle's pileup of the same reads (positions, depths, strand counts, frequencies,
tensors: bit-exact).  generate_SNP_pileups.py:156-186 via the column contract of SURVEY.md appendix C."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

L = 640_000
DCT = dict(threshold=[0.3, 0.7], mincov=2, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont", supplementary=False)
NIB = {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15}
OPS = {"M": 0, "I": 1, "D": 2, "N": 3, "S": 4, "H": 5, "P": 6, "=": 7, "X": 8}


def _read(rng, ref, start, span, style):
    """-> (cigar [(op, len)], query nibbles) of a read covering ref[start : start + span]"""
    alt = {65: b"CGT", 67: b"AGT", 71: b"ACT", 84: b"ACG", 78: b"ACGT"}          # reference N: any base
    cig, q = [], []
    p, end = start, start + span

    def bases(n):
        return [int(x) for x in rng.choice([1, 2, 4, 8], n)]

    def match(n, eqx):
        nonlocal p
        seg = ref[p:p + n]
        p += n
        if not eqx:
            out = [NIB[chr(c)] for c in seg]
            for k in np.nonzero(rng.random(n) < 0.04)[0]:
                out[k] = NIB[chr(rng.choice(list(alt[int(seg[k])])))]
            cig.append(("M", n)); q.extend(out)
        else:                                   # '=' runs and single 'X'
            k0 = 0
            mism = np.nonzero(rng.random(n) < 0.04)[0]
            for k in mism:
                if k > k0:
                    cig.append(("=", int(k - k0))); q.extend(NIB[chr(c)] for c in seg[k0:k])
                cig.append(("X", 1)); q.append(NIB[chr(rng.choice(list(alt[int(seg[k])])))])
                k0 = k + 1
            if n > k0:
                cig.append(("=", n - k0)); q.extend(NIB[chr(c)] for c in seg[k0:n])

    if style == "clips":
        cig.append(("H", 7)); cig.append(("S", 11)); q.extend(bases(11))
    while p < end:
        room = end - p
        if style == "dense":                    # three operations per two reference positions
            match(1, False)
            cig.append(("I", 1)); q.extend(bases(1))
            if p < end:
                cig.append(("D", 1)); p += 1
            continue
        mean = 40 if style == "long" else 14
        match(int(min(room, 1 + rng.geometric(1.0 / mean))), style == "clips")
        if p >= end:
            break
        r = rng.random()
        if r < 0.4:
            n = int(rng.integers(1, 6)); cig.append(("I", n)); q.extend(bases(n))
        elif r < 0.8:
            n = int(min(end - p, rng.integers(1, 9))); cig.append(("D", n)); p += n
        elif r < 0.9 and style == "clips":
            n = int(min(end - p, rng.integers(50, 700))); cig.append(("N", n)); p += n
        elif style == "clips":
            cig.append(("P", 2))
    if cig[-1][0] in ("D", "N"):                # a read ends on a base
        cig.pop()
    if style == "clips":
        cig.append(("S", 5)); q.extend(bases(5)); cig.append(("H", 3))
    return cig, q


def build_reads():
    from nanocaller_b200.host.readset import ReadSet
    rng = np.random.default_rng(7)
    ref = rng.choice(np.frombuffer(b"ACGT", np.uint8), L)
    ref[1000:1010] = ord("N")
    plan = []
    for k in range(260):                        # ordinary reads, every start offset modulo 8
        plan.append((int(rng.integers(0, 60_000)) * 8 // 8 + k % 8 + 8 * int(rng.integers(0, 40)), int(rng.integers(300, 9000)), "ont"))
    for k in range(24):
        plan.append((2000 + 37 * k, int(rng.integers(600, 2500)), "dense"))
    for k in range(30):
        plan.append((int(rng.integers(0, 200_000)), int(rng.integers(500, 6000)), "clips"))
    for k in range(3):
        plan.append((3 + 5 * k, 560_000 + 1000 * k, "long"))      # > 1024 blocks of 512 positions
    plan.sort()
    pos, flag, cig_off, cigar, seq_off, l_seq, seq4 = [], [], [0], [], [0], [], []
    for i, (st, span, style) in enumerate(plan):
        span = min(span, L - st - 1)
        cg, q = _read(rng, ref, st, span, style)
        if i % 41 == 5:                          # sequence shorter than the CIGAR claims: the missing bases read as '*'
            q = q[:max(1, len(q) - 60)]
        pos.append(st); flag.append(16 if rng.random() < 0.5 else 0)
        cigar.extend((n << 4) | OPS[o] for o, n in cg); cig_off.append(len(cigar))
        nib = np.asarray(q + [0] * (len(q) & 1), np.uint8)
        seq4.extend(((nib[0::2] << 4) | nib[1::2]).tolist()); seq_off.append(len(seq4)); l_seq.append(len(q))
    return ReadSet("chrE", ref, pos, flag, cig_off, np.asarray(cigar, np.uint32), seq_off, l_seq, np.asarray(seq4, np.uint8))


@pytest.fixture(scope="module")
def reads():
    return build_reads()


def test_row_fill_special_paths_match_oracle(reads):
    from nanocaller_b200.host import snp_pileups
    from oracle import snp_oracle as O
    rs = reads
    assert (np.diff(rs.cigar_off).max() > 20_000) and (rs.ref_end - rs.pos).max() > 1024 * 512
    chunks = [{"chrom": "chrE", "start": 1, "end": 70_000, "ploidy": "diploid"},
              {"chrom": "chrE", "start": 180_001, "end": 215_000, "ploidy": "diploid"},
              {"chrom": "chrE", "start": 540_001, "end": 566_000, "ploidy": "diploid"}]
    ctx = snp_pileups.context(0)
    snp_pileups._staged.clear()
    snp_pileups.scan_chunks(ctx, rs, DCT, chunks, "diploid")
    mat, meta, depth, count = ctx.snp_fetch()
    got = snp_pileups.unpack(mat, meta, depth, count, len(chunks))
    total = 0
    for g, ch in zip(got, chunks):
        pos, ref, wmat, dp, freq, wdepth, fwd, rev = O.get_snp_testing_candidates(rs, DCT, ch)
        assert len(pos) == len(g[0])
        total += len(pos)
        if len(pos) == 0:
            continue
        np.testing.assert_array_equal(np.asarray(g[0], np.int64), np.asarray(pos, np.int64))
        np.testing.assert_array_equal(np.asarray(g[2]).astype(np.int16), np.asarray(wmat).astype(np.int16))
        np.testing.assert_array_equal(np.asarray(g[3], np.int64), np.asarray(dp, np.int64))
        np.testing.assert_array_equal(np.asarray(g[4], np.float64), np.asarray(freq, np.float64))
        np.testing.assert_array_equal(np.asarray(g[6], np.int64), np.asarray(fwd, np.int64))
        np.testing.assert_array_equal(np.asarray(g[7], np.int64), np.asarray(rev, np.int64))
        assert float(g[5]) == float(wdepth)
    assert total > 500
