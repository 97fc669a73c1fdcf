"""CPU: the library's host-side affine alignment (nc_nw_trace, stands in for parasail) against its numpy statement
in oracle/star_msa.py, and the host allele_prediction against the oracle's line-by-line port."""
import numpy as np

from nanocaller_b200.host import capi, indel_pileups
from oracle import indel_oracle, star_msa


def _rand_pair(rng):
    m = rng.randint(1, 200)
    ref = "".join(rng.choice(list("AGTC"), m))
    q = list(ref)
    for _ in range(rng.randint(0, 6)):
        k = rng.randint(0, max(1, len(q)))
        r = rng.rand()
        if r < 0.4 and q:
            del q[k:k + rng.randint(1, 12)]
        elif r < 0.8:
            q[k:k] = list(rng.choice(list("AGTC"), rng.randint(1, 12)))
        elif q:
            q[min(k, len(q) - 1)] = rng.choice(list("AGTC"))
    return "".join(q), ref


def test_nw_trace_matches_numpy_statement():
    rng = np.random.RandomState(3)
    for _ in range(150):
        q, ref = _rand_pair(rng)
        want = [(star_msa.CIGAR_CODE[o], l) for o, l in star_msa.nw_trace(q, ref, 9, 1, 20, -10)]
        got = capi.nw_trace(star_msa.encode(q), star_msa.encode(ref), 9, 1, 20, -10)
        assert got == want, (q, ref)
    assert capi.nw_trace(np.zeros(0, np.uint8), star_msa.encode("ACG")) == [(2, 3)]
    assert capi.nw_trace(star_msa.encode("ACG"), np.zeros(0, np.uint8)) == [(1, 3)]


def test_allele_prediction_matches_oracle():
    rng = np.random.RandomState(4)
    for _ in range(80):
        q, ref = _rand_pair(rng)
        if not q:
            continue
        for mr in (10, 40):
            assert indel_pileups.allele_prediction(q, ref, mr) == indel_oracle.allele_prediction(q, ref, mr)


def test_batched_allele_prediction_matches_the_python_walk():
    """nc_allele_predict_batch (C++, threaded) vs host allele_prediction (the reference's control flow in Python), item by item."""
    rng = np.random.RandomState(9)
    pairs = []
    while len(pairs) < 300:
        q, ref = _rand_pair(rng)
        if q:
            pairs.append((q, ref, int(rng.choice([10, 40, 50]))))
    pairs.append(("ACGTACGTAC", "ACGTACGTAC", 0))          # identical: no indel within range, no mismatch -> (None, None)
    pairs.append(("ACGTTCGTACGGA", "ACGTACGTACGGA", 0))    # mismatch only
    alt = [star_msa.encode(q) for q, _, _ in pairs]
    ref = [star_msa.encode(r) for _, r, _ in pairs]
    alt_off = np.concatenate([[0], np.cumsum([len(a) for a in alt])[:-1]])
    ref_off = np.concatenate([[0], np.cumsum([len(r) for r in ref])[:-1]])
    for threads in (1, 4):
        ro, ao = capi.allele_predict_batch(np.concatenate(alt), alt_off, [len(a) for a in alt], np.concatenate(ref), ref_off,
                                           [len(r) for r in ref], [m for _, _, m in pairs], threads=threads)
        for k, (q, r, m) in enumerate(pairs):
            want = indel_pileups.allele_prediction(q, r, m)
            got = (None, None) if ro[k] < 0 else (r[:ro[k]], q[:ao[k]])
            assert got == want, (q, r, m, got, want)
    assert capi.allele_predict_batch(np.zeros(0, np.uint8), [], [], np.zeros(0, np.uint8), [], [], [])[0].size == 0
