"""Helpers to load golden fixtures (tests/golden/*.npz) and replay their inputs."""
import json
import os

import numpy as np

from tests.golden.cases import CASES, case_inputs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def load_case(name):
    """-> (readset, dct, chunks, bed, golden npz) with the input checksum verified."""
    if name not in _cache:
        g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        rs, dct, regions, cpu, bed = case_inputs(name)
        assert rs.checksum() == str(g["input_checksum"]), (
            "synthetic inputs of golden case %r drifted; regenerate with tests/golden/make_golden.py" % name)
        chunks = json.loads(str(g["chunks_json"]))
        _cache[name] = (rs, dct, chunks, bed, g)
    return _cache[name]


def golden_chunk(g, ci):
    return {k: g["c%d_%s" % (ci, k)] for k in ("pos", "ref", "mat", "dp", "freq", "depth", "fwd", "rev")}


def available_cases():
    return [n for n in CASES if os.path.exists(os.path.join(GOLDEN_DIR, n + ".npz"))]


def check_tensor_invariants(x, ref_code, n_left, n_right, sample_depth, dp, acgt_depth, maxcov):
    """Size-independent properties of the SNP pileup tensors (consequences of generate_SNP_pileups.py:200-263).
    x int [n,5,41,5]; ref_code [n] (A0 G1 T2 C3); n_left / n_right = real columns either side of column 20;
    acgt_depth = reads with an A/G/T/C base at the candidate before down-sampling."""
    import numpy as np
    n = len(x)
    row0 = x[:, 0]                                                              # [n,41,5]
    assert set(np.unique(row0[..., :4]).tolist()) <= {0, 1} and (row0[..., :4].sum(-1) <= 1).all() and (row0[..., 4] == 0).all()
    cols = np.arange(41)[None, :]
    nl, nr = np.asarray(n_left, int)[:, None], np.asarray(n_right, int)[:, None]
    real = (cols == 20) | ((cols < 20) & (cols >= 20 - nl)) | ((cols > 20) & (cols <= 20 + nr))
    assert (x.transpose(0, 2, 1, 3)[~real] == 0).all()                          # padded columns are all-zero in all 5 rows (:254)
    assert (row0[np.arange(n), 20, np.asarray(ref_code, int)] == 1).all()        # column 20 is the candidate
    ch4 = x[:, 1:, :, 4]                                                        # channel 4 marks the reference row on real columns (:252)
    want4 = (np.arange(4)[None, :, None] == np.asarray(ref_code, int)[:, None, None]) & real[:, None, :]
    assert (ch4 == want4).all()
    cnt = x[:, 1:, :, :4]                                                       # [n,4,41,4]
    ref_col = np.broadcast_to(row0[..., :4].astype(bool)[:, None], cnt.shape)
    assert (cnt[ref_col] <= 0).all() and (cnt[~ref_col] >= 0).all()             # counts negated exactly at the column's reference base (:253)
    c20 = np.abs(cnt[:, :, 20, :]).copy()
    diag = c20[:, np.arange(4), np.arange(4)].sum(1)
    assert (c20.sum((1, 2)) == diag).all()                                      # at the candidate column mat[i,20,b] = 0 for i != b
    no_down = np.asarray(dp) <= maxcov
    assert (diag[no_down] == np.asarray(acgt_depth)[no_down]).all()             # ... and it counts the sampled A/G/T/C reads
    assert (np.abs(cnt).sum((1, 3)) <= np.asarray(sample_depth)[:, None]).all() # every column is bounded by the sampled depth
