"""Pins oracle/indel_oracle.py against fixtures produced by the UNMODIFIED reference generate_indel_pileups.py run over
oracle/shim (tests/golden/make_golden_indel.py): candidate positions, the three [N,5,128,2] tensors, allele strings and
phase sets.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import indel_oracle as O
from tests.golden.indel_cases import INDEL_CASES, indel_case_inputs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [n for n in INDEL_CASES if os.path.exists(os.path.join(GOLDEN_DIR, n + ".npz"))]
_cache = {}


def load_indel_case(name):
    if name not in _cache:
        g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        rs, dct, regions, cpu, mcs = indel_case_inputs(name)
        assert rs.checksum() == str(g["input_checksum"])
        _cache[name] = (rs, dct, json.loads(str(g["chunks_json"])), g)
    return _cache[name]


@pytest.mark.parametrize("name", CASES)
def test_indel_oracle_matches_reference(name):
    rs, dct, chunks, g = load_indel_case(name)
    for ci, chunk in enumerate(chunks):
        if chunk["ploidy"] == "haploid":
            pos, x2, alleles = O.get_indel_testing_candidates_haploid(rs, dct, chunk)
            x0 = x1 = x2
            phase = []
        else:
            pos, x0, x1, x2, alleles, phase = O.get_indel_testing_candidates(rs, dct, chunk)
        want_pos = g["c%d_pos" % ci]
        assert list(pos) == list(want_pos), (name, ci)
        if len(want_pos) == 0:
            continue
        for k, x in (("x0", x0), ("x1", x1), ("x2", x2)):
            np.testing.assert_array_equal(np.asarray(x, np.float64), g["c%d_%s" % (ci, k)].astype(np.float64), err_msg="%s %d %s" % (name, ci, k))
        want_alleles = json.loads(str(g["c%d_alleles" % ci]))
        assert json.loads(json.dumps(alleles)) == want_alleles
        assert list(phase) == json.loads(str(g["c%d_phase" % ci]))


def test_get_chunks_indel_grid():
    from oracle.snp_oracle import get_chunks
    rs, dct, chunks, g = load_indel_case("indel_ont")
    assert get_chunks([("chr20", 1, 60_000, "diploid")], 2, max_chunk_size=100000) == chunks
