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
