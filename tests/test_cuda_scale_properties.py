"""GPU, at bench scale (a 20 Mb contig of the configs[1] generator: ~195 k candidate sites, 0.59 G aligned bases — too big
for the oracle): size-independent properties of the tensors (SURVEY.md 8a invariants, all consequences of
generate_SNP_pileups.py:200-263), idempotence of the whole path, agreement of the tensor-core and the fp32 CNN paths, and
exact agreement of a sub-region with the oracle (the oracle finishes a 60 kb chunk in seconds)."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

L = 20_000_000
DCT = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont", supplementary=False)


from tests.golden_util import check_tensor_invariants  # noqa: E402


@pytest.fixture(scope="module")
def world():
    from nanocaller_b200.synth import make_world
    from oracle.snp_oracle import get_chunks
    rs = make_world(chrom="chr20", preset="ont", contig_len=L, seed=20, coverage=30.0).reads
    return rs, get_chunks([("chr20", 1, L, "diploid")], 1)


def _run(ctx, rs, chunks, impl):
    from nanocaller_b200.host import capi
    ctx.stage_reads(rs)
    n = ctx.snp_scan(capi.snp_params(DCT, "diploid"), [(c["start"], c["end"]) for c in chunks])
    probs = ctx.snp_forward(normalize=True, impl=impl)
    mat, meta, depth, count = ctx.snp_fetch()
    return n, mat, meta, depth, count, probs


def test_tensor_invariants_idempotence_and_cnn_paths(world):
    from nanocaller_b200.host import capi, weights as W
    rs, chunks = world
    ctx = capi.Context(0)
    tensors, m = W.load_model("snp", "ONT-HG002")
    ctx.load_snp_weights(W.pack_snp_blob(tensors, False), m["train_coverage"], False)
    n, mat, meta, depth, count, probs = _run(ctx, rs, chunks, 0)
    assert n > 150_000 and len(chunks) == 40 and int(count.sum()) == n
    x = mat[:, :1025].reshape(n, 5, 41, 5).astype(np.int32)
    assert (mat[:, 1025:] == 0).all()                                            # row padding
    # ---- order: (chunk, pos) ascending; a site on a shared chunk boundary appears in both chunks (utils.py:79-80)
    key = meta["chunk"].astype(np.int64) * (1 << 32) + meta["pos"]
    assert (np.diff(key) > 0).all()
    starts = np.array([c["start"] for c in chunks]); ends = np.array([c["end"] for c in chunks])
    assert (meta["pos"] >= starts[meta["chunk"]]).all() and (meta["pos"] <= ends[meta["chunk"]]).all()
    dup = np.isin(meta["pos"], ends[:-1])
    assert dup.sum() % 2 == 0 and (np.sort(meta["pos"][dup])[0::2] == np.sort(meta["pos"][dup])[1::2]).all()
    check_tensor_invariants(x, meta["ref_code"], meta["n_left"], meta["n_right"], meta["sample_depth"], meta["dp"],
                            meta["fwd"].astype(np.int64).sum(1) + meta["rev"].astype(np.int64).sum(1), DCT["maxcov"])
    assert (meta["dp"] >= DCT["mincov"]).all() and (meta["alt"] >= DCT["min_allele_freq"] * meta["dp"] - 1e-9).all()
    # chunk depth = mean sampled depth of the chunk's candidates (:274)
    for c in (0, 17, 39):
        sel = meta["chunk"] == c
        assert depth[c] == meta["sample_depth"][sel].astype(np.int64).sum() / sel.sum()
    # ---- probabilities: finite, in [0,1]; tensor-core and fp32 CUDA-core paths agree within the tolerance
    assert np.isfinite(probs).all() and probs.min() >= 0 and probs.max() <= 1
    probs1 = ctx.snp_forward(normalize=True, impl=1)
    assert np.abs(probs - probs1).max() < 1e-4
    # ---- idempotence: the whole path again gives bit-identical tensors, metadata and probabilities
    crc = (zlib.crc32(mat.tobytes()), zlib.crc32(meta.tobytes()), zlib.crc32(probs.tobytes()))
    n2, mat2, meta2, depth2, count2, probs2 = _run(ctx, rs, chunks, 0)
    assert n2 == n and (zlib.crc32(mat2.tobytes()), zlib.crc32(meta2.tobytes()), zlib.crc32(probs2.tobytes())) == crc
    # ---- a different chunk grid (--cpu 7) moves chunk boundaries but not the per-site tensors of interior sites
    from oracle.snp_oracle import get_chunks
    ch7 = get_chunks([("chr20", 1, L, "diploid")], 50)
    assert len(ch7) != len(chunks)
    n7 = ctx.snp_scan(capi.snp_params(DCT, "diploid"), [(c["start"], c["end"]) for c in ch7])
    mat7, meta7, _, _ = ctx.snp_fetch()
    # sites at least 50 kb away from every boundary of both grids see the same pileup window
    b = np.unique(np.concatenate([starts, ends, [c["start"] for c in ch7], [c["end"] for c in ch7]]))
    def interior(p):
        i = np.searchsorted(b, p)
        lo = b[np.clip(i - 1, 0, len(b) - 1)]; hi = b[np.clip(i, 0, len(b) - 1)]
        return (p - lo > 50_000) & (hi - p > 50_000)
    a = interior(meta["pos"]); a7 = interior(meta7["pos"])
    assert a.sum() == a7.sum() > 10_000 and (meta["pos"][a] == meta7["pos"][a7]).all()
    assert (mat[a] == mat7[a7]).all()
    ctx.close()


def test_subregion_of_the_big_contig_matches_oracle(world):
    """One 60 kb chunk in the middle of the 20 Mb contig (reads and neighbours come from the full read set): bit-exact vs the oracle."""
    from nanocaller_b200.host import snp_pileups
    from oracle import snp_oracle as O
    rs, _ = world
    chunk = {"chrom": "chr20", "start": 9_970_001, "end": 10_030_000, "ploidy": "diploid"}
    ctx = snp_pileups.context(0)
    snp_pileups._staged.clear()
    snp_pileups.scan_chunks(ctx, rs, DCT, [chunk], "diploid")
    mat, meta, depth, count = ctx.snp_fetch()
    got = snp_pileups.unpack(mat, meta, depth, count, 1)[0]
    pos, ref, wmat, dp, freq, wdepth, fwd, rev = O.get_snp_testing_candidates(rs, DCT, chunk)
    assert len(pos) == len(got[0]) > 300
    np.testing.assert_array_equal(np.asarray(got[0], np.int64), np.asarray(pos, np.int64))
    np.testing.assert_array_equal(np.asarray(got[2]).astype(np.int16), np.asarray(wmat).astype(np.int16))
    np.testing.assert_array_equal(np.asarray(got[3], np.int64), np.asarray(dp, np.int64))
    np.testing.assert_array_equal(np.asarray(got[4], np.float64), np.asarray(freq, np.float64))
    assert float(got[5]) == float(wdepth)
