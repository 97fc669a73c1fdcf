"""ORACLE / TEST INFRASTRUCTURE ONLY — stand-in for the `parasail` module (not installable here) so the reference's
generate_indel_pileups.py imports and runs UNCHANGED.  Implements only what the reference touches
(generate_indel_pileups.py:10,79-80): `matrix_create` and `nw_trace(...).cigar.seq`, backed by the repo's own
affine-gap alignment defined in oracle/star_msa.py (parity with real parasail tie-breaking is unpinned)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import star_msa  # noqa: E402


class _Matrix:
    def __init__(self, alphabet, match, mismatch):
        self.alphabet, self.match, self.mismatch = alphabet, match, mismatch


def matrix_create(alphabet, match, mismatch):
    return _Matrix(alphabet, match, mismatch)


class _Cigar:
    def __init__(self, seq):
        self.seq = seq


class _Result:
    def __init__(self, seq):
        self.cigar = _Cigar(seq)


def nw_trace(s1, s2, gap_open, gap_extend, matrix):
    ops = star_msa.nw_trace(s1, s2, gap_open, gap_extend, matrix.match, matrix.mismatch)
    return _Result(star_msa.cigar_words(ops))
