"""CPU: the host genotype/record code (host/snp_caller.records_from_calls) produces the same VCF lines as the
restatement of snpCaller.py:113-163 / :183-198 in oracle/snp_caller_oracle.py, including every branch of the
decision tree and ties."""
import numpy as np

from nanocaller_b200.host.snp_caller import records_from_calls
from oracle import snp_caller_oracle as O


def _inputs(n, seed):
    rng = np.random.RandomState(seed)
    probs = rng.rand(n, 4).astype(np.float32)
    # force every branch: 0, 1, 2, 3 alleles over 0.5; ties; exact 0.5; near-1 (QUAL cap)
    probs[0] = [0.1, 0.2, 0.3, 0.4]
    probs[1] = [0.9, 0.2, 0.1, 0.0]
    probs[2] = [0.9, 0.8, 0.1, 0.0]
    probs[3] = [0.9, 0.8, 0.7, 0.0]
    probs[4] = [0.5, 0.5, 0.1, 0.1]
    probs[5] = [1.0, 0.0, 0.0, 0.0]
    probs[6] = [0.7, 0.7, 0.7, 0.7]
    probs[7] = [0.49999997, 0.5, 0.2, 0.2]
    ref_code = rng.randint(0, 4, n)
    ref_code[:8] = [0, 0, 0, 3, 1, 1, 2, 0]
    onehot = np.eye(4, dtype=np.int32)[ref_code]
    dp = rng.randint(4, 120, n)
    fwd = rng.randint(0, 40, (n, 4)).astype(np.float64)
    rev = rng.randint(0, 40, (n, 4)).astype(np.float64)
    alt = rng.randint(1, 50, n)
    freq = alt / np.maximum(dp, alt).astype(np.float64)
    pos = np.sort(rng.randint(1, 10 ** 7, n))
    return pos, ref_code, onehot, probs, dp, freq, fwd, rev


def test_diploid_records_match_oracle():
    pos, rc, onehot, probs, dp, freq, fwd, rev = _inputs(3000, 1)
    got = records_from_calls("chr20", pos, rc, probs, dp, freq, fwd, rev, "diploid")
    want = O.diploid_records("chr20", pos, onehot, probs, dp, freq, fwd, rev)
    assert len(got) == len(want) == 3000            # every candidate yields a record
    assert got == want
    kinds = {ln.split("\t")[6] for ln in got}
    assert kinds == {"PASS", "REF", "LOW"}
    gts = {ln.rstrip("\n").split("\t")[9].split(":")[0] for ln in got}
    assert {"0/1", "1/1", "1/2", "./."} <= gts


def test_haploid_records_match_oracle():
    pos, rc, onehot, probs, dp, freq, fwd, rev = _inputs(2000, 2)
    probs = probs / probs.sum(1, keepdims=True)
    got = records_from_calls("chrY", pos, rc, probs, dp, freq, fwd, rev, "haploid")
    want = O.haploid_records("chrY", pos, onehot, probs, dp, freq)
    assert got == want


def test_empty():
    assert records_from_calls("c", [], [], np.zeros((0, 4), np.float32), [], [], [], []) == []


def test_native_record_formatter_matches_python(monkeypatch=None):
    """nc_format_snp_records (C++, threaded) vs records_from_calls (numpy / Python), line for line, on random call records with
    probabilities pushed onto the decision boundaries."""
    from nanocaller_b200.host import capi, snp_caller
    rng = np.random.RandomState(12)
    n = 20_000
    probs = rng.rand(n, 4).astype(np.float32)
    probs[rng.rand(n, 4) < 0.08] = np.float32(0.5)                       # exactly on the threshold
    probs[rng.rand(n, 4) < 0.05] = np.float32(1.0)
    probs[rng.rand(n, 4) < 0.05] = np.float32(0.0)
    tie = rng.rand(n) < 0.1
    probs[tie, 1] = probs[tie, 3]                                         # ties: the stable argsort decides
    pos = np.sort(rng.randint(1, 5_000_000, n)).astype(np.int32)
    ref = rng.randint(0, 4, n).astype(np.uint8)
    fwd = rng.randint(0, 90, (n, 4)).astype(np.uint16)
    rev = rng.randint(0, 90, (n, 4)).astype(np.uint16)
    dp = (fwd.sum(1) + rev.sum(1) + rng.randint(1, 9, n)).astype(np.int32)
    alt = np.array([max(int(fwd[i, b]) + int(rev[i, b]) for b in range(4) if b != ref[i]) for i in range(n)], np.int32)
    freq = alt.astype(np.float64) / dp.astype(np.float64)
    for haploid in (False, True):
        want = snp_caller.records_from_calls("chr7", pos, ref, probs, dp, freq, fwd, rev, "haploid" if haploid else "diploid")
        for threads in (1, 5):
            blob, off, ok = capi.format_snp_records("chr7", pos, ref, probs, dp, alt, fwd, rev, haploid=haploid, threads=threads)
            got = [blob[off[i]:off[i + 1]].decode() for i in range(n) if off[i + 1] > off[i]]
            assert got == want
            flt = [ln.split("\t")[6] == "PASS" for ln in want]
            assert ok[np.diff(off) > 0].tolist() == flt
    blob, off, ok = capi.format_snp_records("c", [], [], np.zeros((0, 4), np.float32), [], [], np.zeros((0, 4)), np.zeros((0, 4)))
    assert blob == b"" and off.tolist() == [0]
