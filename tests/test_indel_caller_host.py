"""CPU: host indel genotype/record code (G2) vs the line-by-line restatement of indelCaller.py:88-152,173-179."""
import numpy as np

from nanocaller_b200.host.indel_caller import haploid_records_from_calls, records_from_calls
from oracle import indel_caller_oracle as O


def _alleles(rng, n):
    def one():
        r = rng.rand()
        if r < 0.25:
            return (None, None)
        ref = "".join(rng.choice(list("AGTC"), rng.randint(1, 12)))
        alt = "".join(rng.choice(list("AGTC"), rng.randint(1, 12)))
        return (ref, alt)
    out = []
    for _ in range(n):
        a0, a1, at = one(), one(), one()
        if rng.rand() < 0.2:
            a1 = a0
        out.append([a0, a1, at])
    return out


def test_diploid_records_match_oracle():
    rng = np.random.RandomState(5)
    n = 2500
    probs = rng.dirichlet([0.4] * 4, n).astype(np.float32)
    probs[:20, 0] = 0.97
    pos = np.cumsum(rng.randint(1, 30, n))
    alleles = _alleles(rng, n)
    phase = [int(x) if x else None for x in rng.randint(0, 3, n)]
    got = records_from_calls("chr1", pos, probs, alleles, phase)
    want = O.diploid_records("chr1", pos, probs, alleles, phase)
    assert got == want and len(got) > 300
    gts = {ln.rstrip("\n").split("\t")[9].split(":")[0] for ln in got}
    assert gts == {"1/1", "1|2", "0|1", "1|0"}


def test_haploid_records_match_oracle():
    rng = np.random.RandomState(6)
    n = 1500
    probs = rng.rand(n, 1).astype(np.float32)
    pos = np.cumsum(rng.randint(1, 30, n))
    alleles = [a[2] for a in _alleles(rng, n)]
    assert haploid_records_from_calls("chrX", pos, probs, alleles) == O.haploid_records("chrX", pos, probs, alleles)
