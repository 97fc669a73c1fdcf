"""Pins the numpy oracle (oracle/snp_oracle.py) against golden fixtures produced by the UNMODIFIED
reference modules (generate_SNP_pileups.py, utils.py) imported over oracle/shim — see
tests/golden/make_golden.py.  CPU only."""
import json

import numpy as np
import pytest

from oracle import snp_oracle as O
from tests.golden_util import available_cases, golden_chunk, load_case

CASES = available_cases()


def _bed_for(bed, chrom):
    if bed is None:
        return None
    return bed.get(chrom)


@pytest.mark.parametrize("name", CASES)
def test_chunks_match_reference_get_chunks(name):
    rs, dct, chunks, bed, g = load_case(name)
    from tests.golden.cases import CASES as DEF
    _, _, regions, cpu, _ = DEF[name]
    assert O.get_chunks(regions, cpu) == chunks


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_tensors(name):
    rs, dct, chunks, bed, g = load_case(name)
    for ci, chunk in enumerate(chunks):
        want = golden_chunk(g, ci)
        pos, ref, mat, dp, freq, depth, fwd, rev = O.get_snp_testing_candidates(rs, dct, chunk, _bed_for(bed, chunk["chrom"]))
        n = len(want["pos"])
        assert len(pos) == n, (name, ci)
        if n == 0:
            assert depth == 0
            continue
        np.testing.assert_array_equal(np.asarray(pos, np.int64), want["pos"])
        np.testing.assert_array_equal(np.asarray(ref, np.int8), want["ref"])
        np.testing.assert_array_equal(np.asarray(mat).astype(np.int16), want["mat"])
        np.testing.assert_array_equal(np.asarray(dp, np.int64), want["dp"])
        np.testing.assert_array_equal(np.asarray(freq, np.float64), want["freq"])   # bit-exact float64
        assert float(depth) == float(want["depth"])
        np.testing.assert_array_equal(np.asarray(fwd).astype(np.int16), want["fwd"])
        np.testing.assert_array_equal(np.asarray(rev).astype(np.int16), want["rev"])


@pytest.mark.parametrize("name", [c for c in CASES if c != "empty"])
def test_tensor_invariants(name):
    """SURVEY.md §8a invariants of T1, checked on the reference's own output."""
    rs, dct, chunks, bed, g = load_case(name)
    for ci in range(len(chunks)):
        w = golden_chunk(g, ci)
        m = w["mat"].astype(np.int64)
        if len(m) == 0:
            continue
        real = np.abs(m).sum(axis=(1, 3)) > 0                      # [N,41] columns that are not padding
        assert real[:, 20].all()                                   # candidate sits in column 20
        assert (m[:, 0, :, :4].sum(-1) <= 1).all() and (m[:, 0, :, 4] == 0).all()
        ch4 = m[:, 1:, :, 4]                                       # [N,4,41]
        assert ((ch4.sum(1) == 1) == real).all()
        refrow = w["ref"].argmax(1)
        assert (ch4[np.arange(len(m)), refrow, 20] == 1).all()
        onehot = m[:, 0, :, :4]                                    # [N,41,4]
        body = m[:, 1:, :, :4]                                     # [N,4,41,4]
        assert (body[np.broadcast_to(onehot[:, None] == 1, body.shape)] <= 0).all()
        assert (body[np.broadcast_to(onehot[:, None] == 0, body.shape)] >= 0).all()
        for i in range(4):
            for b in range(4):
                if i != b:
                    assert (m[:, 1 + i, 20, b] == 0).all()


def test_get_cnd_pos_against_reference_source_rules():
    """Spot-check of the distance-bin table on a synthetic neighbour array (ont: nearest bin keeps the
    two FARTHEST sites, generate_SNP_pileups.py:10,16)."""
    nbr = np.arange(1000, 120000, 500)
    l, r = O.get_cnd_pos(60000, nbr, "ont")
    assert l[-2:] == [58000, 58500]            # first two of [58000, 60000)
    assert 59500 not in l and 59000 not in l
    assert r[:2] == [61500, 62000]             # last two of (60000, 62000]
    assert len(l) == 2 + 3 + 4 + 5 + 6 and len(r) == 20


def test_golden_tensors_obey_the_size_independent_invariants():
    """The invariants the GPU scale test relies on (tests/test_cuda_scale_properties.py) hold for the reference's own output."""
    from tests.golden_util import check_tensor_invariants, golden_chunk, load_case
    checked = 0
    for name in ("ont_diploid", "hifi_pacbio", "ul_ont", "short_ont", "lowcov"):
        rs, dct, chunks, bed, g = load_case(name)
        for ci in range(len(chunks)):
            w = golden_chunk(g, ci)
            n = len(w["pos"])
            if n == 0:
                continue
            x = np.asarray(w["mat"]).astype(np.int32).reshape(n, 5, 41, 5)
            ref_code = np.asarray(w["ref"]).argmax(1)
            real = (np.abs(x).sum((1, 3)) > 0)
            n_left = real[:, :20].sum(1); n_right = real[:, 21:].sum(1)
            sampled = np.minimum(np.asarray(w["dp"]), dct["maxcov"])
            acgt = np.asarray(w["fwd"]).sum(1) + np.asarray(w["rev"]).sum(1)
            check_tensor_invariants(x, ref_code, n_left, n_right, sampled, w["dp"], acgt, dct["maxcov"])
            checked += n
    assert checked > 1000


def test_cnn_restatement_equals_the_reference_model_classes():
    """tests/golden/model_probs.npz = outputs of the UNMODIFIED model_architect*.py classes run over oracle/shim/tensorflow (which
    supplies only conv / dense / selu / softmax / sigmoid) with the released weights (tests/golden/make_golden_model_probs.py): the
    restatement the CUDA kernels are held to has the same wiring — all five SNP heads including the unused GT head."""
    import os
    import sys
    from nanocaller_b200.host import weights as W
    from oracle import cnn_oracle
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gdir, "model_probs.npz"))
    sys.path.insert(0, gdir)
    try:
        from make_golden_model_probs import inputs
    finally:
        sys.path.remove(gdir)
    x, ref, xi = inputs(int(g["seed"]), int(g["n"]))
    snp, _ = W.load_model("snp", "ONT-HG002")
    outs = cnn_oracle.snp_model(snp, x, ref)
    for k, o in zip(("snp_A", "snp_G", "snp_T", "snp_C", "snp_GT"), outs):
        assert np.abs(o - g[k]).max() < 2e-5, k
    hap, _ = W.load_model("snp", "haploid")
    assert np.abs(cnn_oracle.haploid_snp_model(hap, x, ref) - g["snp_haploid"]).max() < 2e-5
    ind, _ = W.load_model("indel", "ONT-HG002")
    assert np.abs(cnn_oracle.indel_model(ind, xi) - g["indel"]).max() < 2e-5
    hind, _ = W.load_model("indel", "haploid")
    assert np.abs(cnn_oracle.haploid_indel_model(hind, xi[:, 10:15]) - g["indel_haploid"]).max() < 2e-5
    assert g["snp_A"].std() > 0.05 and g["indel"].std() > 0.05          # the inputs exercise the models
