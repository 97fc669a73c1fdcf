"""GPU parity: CUDA K0/K1/K2 through the C-ABI vs the golden fixtures (unmodified reference) and
vs the numpy oracle on fresh seeded inputs.  Bit-exact: integer tensors, depths, float64 freq."""
import numpy as np
import pytest

from tests.golden_util import available_cases, golden_chunk, load_case

pytestmark = pytest.mark.gpu
CASES = available_cases()


def _ctx():
    from nanocaller_b200.host import snp_pileups
    return snp_pileups.context(0)


def _compare(got, want, tag):
    pos, ref, mat, dp, freq, depth, fwd, rev = got
    n = len(want["pos"])
    assert len(pos) == n, tag
    if n == 0:
        assert depth == 0
        return
    np.testing.assert_array_equal(np.asarray(pos, np.int64), want["pos"], err_msg=str(tag))
    np.testing.assert_array_equal(np.asarray(ref, np.int8), np.asarray(want["ref"], np.int8), err_msg=str(tag))
    np.testing.assert_array_equal(np.asarray(dp, np.int64), want["dp"], err_msg=str(tag))
    np.testing.assert_array_equal(np.asarray(freq, np.float64), want["freq"], err_msg=str(tag))
    np.testing.assert_array_equal(np.asarray(fwd).astype(np.int16), np.asarray(want["fwd"]).astype(np.int16), err_msg=str(tag))
    np.testing.assert_array_equal(np.asarray(rev).astype(np.int16), np.asarray(want["rev"]).astype(np.int16), err_msg=str(tag))
    got_m = np.asarray(mat).astype(np.int16)
    want_m = np.asarray(want["mat"]).astype(np.int16)
    bad = np.nonzero((got_m != want_m).reshape(n, -1).any(1))[0]
    assert len(bad) == 0, "%s: %d/%d tensors differ, first at site %d pos %d" % (tag, len(bad), n, bad[0], pos[bad[0]])
    assert float(depth) == float(want["depth"]), tag


@pytest.mark.parametrize("name", CASES)
def test_drop_in_matches_reference_golden(name):
    """get_snp_testing_candidates(dct, region) chunk by chunk, like snpCaller.py:86."""
    from nanocaller_b200.host import snp_pileups, sources
    rs, dct, chunks, bed, g = load_case(name)
    sources.unregister_all()
    sources.register_source("mem://bam", rs)
    if bed is not None:
        sources.register_bed("mem://bed", bed)
    d = dict(dct, sam_path="mem://bam", fasta_path="mem://bam")
    for ci, chunk in enumerate(chunks):
        got = snp_pileups.get_snp_testing_candidates(d, chunk)
        _compare(got, golden_chunk(g, ci), (name, ci))


@pytest.mark.parametrize("name", CASES)
def test_batched_chunks_match_reference_golden(name):
    """All chunks of a contig in ONE scan call (the product path): same per-chunk results, boundary
    candidates duplicated, per-chunk depth."""
    from nanocaller_b200.host import snp_pileups
    rs, dct, chunks, bed, g = load_case(name)
    ctx = _ctx()
    b = bed.get(chunks[0]["chrom"]) if bed else None
    snp_pileups.scan_chunks(ctx, rs, dct, chunks, chunks[0]["ploidy"], b)
    mat, meta, depth, count = ctx.snp_fetch()
    per_chunk = snp_pileups.unpack(mat, meta, depth, count, len(chunks))
    for ci in range(len(chunks)):
        _compare(per_chunk[ci], golden_chunk(g, ci), (name, ci, "batched"))
        if len(per_chunk[ci][0]):
            assert (meta["chunk"][meta["chunk"] == ci] == ci).all()


def test_matches_oracle_on_fresh_world_with_downsampling():
    """Fresh seed, 200x coverage > maxcov=160: exercises the deterministic first-maxcov sample rule and
    counts > 127 (int8 would overflow)."""
    from nanocaller_b200.host import snp_pileups
    from nanocaller_b200.synth import make_world
    from oracle import snp_oracle as O
    from tests.golden.cases import BASE_DCT
    rs = make_world(chrom="chrZ", preset="ont", contig_len=40_000, seed=91, coverage=200.0).reads
    dct = dict(BASE_DCT)
    chunk = {"chrom": "chrZ", "start": 1, "end": 40_000, "ploidy": "diploid"}
    ctx = _ctx()
    snp_pileups.scan_chunks(ctx, rs, dct, [chunk], "diploid")
    mat, meta, depth, count = ctx.snp_fetch()
    got = snp_pileups.unpack(mat, meta, depth, count, 1)[0]
    want = O.get_snp_testing_candidates(rs, dct, chunk)
    keys = ("pos", "ref", "mat", "dp", "freq", "depth", "fwd", "rev")
    w = dict(zip(keys, want))
    assert np.abs(np.asarray(w["mat"])).max() > 127
    assert meta["sample_depth"].max() == 160 and meta["dp"].max() > 160
    _compare(got, w, "fresh200x")


@pytest.mark.parametrize("maxcov,coverage,seq", [(20, 30.0, "ont"), (160, 30.0, "ul_ont"), (12, 90.0, "pacbio")])
def test_matches_oracle_on_fresh_world_small_maxcov(maxcov, coverage, seq):
    """Down-sampling inside the bit-parallel path (run lists of <= 128 reads, maxcov below the depth) and other sequencing modes."""
    from nanocaller_b200.host import snp_pileups
    from nanocaller_b200.synth import make_world
    from oracle import snp_oracle as O
    from tests.golden.cases import BASE_DCT
    rs = make_world(chrom="chrQ", preset="ont", contig_len=50_000, seed=77, coverage=coverage).reads
    dct = dict(BASE_DCT, maxcov=maxcov, seq=seq)
    chunks = [{"chrom": "chrQ", "start": 1, "end": 30_000, "ploidy": "diploid"}, {"chrom": "chrQ", "start": 30_000, "end": 50_000, "ploidy": "diploid"}]
    ctx = _ctx()
    snp_pileups._staged.clear()
    snp_pileups.scan_chunks(ctx, rs, dct, chunks, "diploid")
    mat, meta, depth, count = ctx.snp_fetch()
    per = snp_pileups.unpack(mat, meta, depth, count, len(chunks))
    keys = ("pos", "ref", "mat", "dp", "freq", "depth", "fwd", "rev")
    for ci, ch in enumerate(chunks):
        w = dict(zip(keys, O.get_snp_testing_candidates(rs, dct, ch)))
        assert len(w["pos"]) > 50
        _compare(per[ci], w, ("fresh", maxcov, seq, ci))
    if maxcov < coverage:
        assert meta["sample_depth"].max() == maxcov and meta["dp"].max() > maxcov


def test_unsorted_and_overlapping_chunks_match_oracle():
    """The batched scan takes chunks as given: out of order, overlapping, nested and tiny ones; every chunk equals the oracle's
    result for that chunk (runs of consecutive slots then straddle jumps in position: the generic per-site path)."""
    from nanocaller_b200.host import snp_pileups
    from nanocaller_b200.synth import make_world
    from oracle import snp_oracle as O
    from tests.golden.cases import BASE_DCT
    rs = make_world(chrom="chrU", preset="ont", contig_len=160_000, seed=78, coverage=25.0).reads
    dct = dict(BASE_DCT)
    spans = [(120_001, 160_000), (1, 40_000), (30_000, 90_000), (35_000, 36_000), (159_990, 160_000), (70_000, 70_000)]
    chunks = [{"chrom": "chrU", "start": a, "end": b, "ploidy": "diploid"} for a, b in spans]
    ctx = _ctx()
    snp_pileups._staged.clear()
    snp_pileups.scan_chunks(ctx, rs, dct, chunks, "diploid")
    mat, meta, depth, count = ctx.snp_fetch()
    per = snp_pileups.unpack(mat, meta, depth, count, len(chunks))
    keys = ("pos", "ref", "mat", "dp", "freq", "depth", "fwd", "rev")
    total = 0
    for ci, ch in enumerate(chunks):
        w = dict(zip(keys, O.get_snp_testing_candidates(rs, dct, ch)))
        _compare(per[ci], w, ("unsorted", ci))
        total += len(w["pos"])
    assert total == len(meta) > 1000


def test_empty_inputs():
    from nanocaller_b200.host import capi, snp_pileups
    from nanocaller_b200.host.readset import ReadSet
    from tests.golden.cases import BASE_DCT
    ctx = _ctx()
    rs = ReadSet.from_records("e", "ACGT" * 50, [])
    snp_pileups.scan_chunks(ctx, rs, dict(BASE_DCT), [{"chrom": "e", "start": 1, "end": 200, "ploidy": "diploid"}], "diploid")
    mat, meta, depth, count = ctx.snp_fetch()
    assert ctx.n_sites == 0 and count.tolist() == [0] and depth.tolist() == [0.0]
    # no chunks at all
    ctx.snp_scan(capi.snp_params(dict(BASE_DCT), "diploid"), [])
    assert ctx.n_sites == 0


def test_error_paths():
    from nanocaller_b200.host import capi
    from tests.golden.cases import BASE_DCT
    ctx = capi.Context(0)
    with pytest.raises(capi.NcError) as e:
        ctx.snp_scan(capi.snp_params(dict(BASE_DCT), "diploid"), [(1, 10)])
    assert e.value.code == capi.NC_ESTATE
    pos = np.array([5, 3], np.int32)
    z2 = np.zeros(2, np.int32)
    with pytest.raises(capi.NcError) as e:
        ctx.stage_arrays(pos, np.zeros(2, np.uint16), np.zeros(3, np.int64), np.zeros(0, np.uint32), np.zeros(3, np.int64),
                         z2, np.zeros(0, np.uint8), np.frombuffer(b"ACGT", np.uint8))
    assert e.value.code == capi.NC_EINVAL
    ctx.close()


def test_blocking_sync_mode_gives_the_same_results():
    """nc_set_blocking_sync only changes how the host waits: same tensors with sleeping waits."""
    from nanocaller_b200.host import capi
    from nanocaller_b200.synth import make_world
    from tests.golden.cases import BASE_DCT
    import zlib
    rs = make_world(chrom="chrW", preset="ont", contig_len=80_000, seed=5, coverage=20.0).reads
    out = []
    for on in (False, True, False):
        ctx = capi.Context(0)
        ctx.set_blocking_sync(on)
        ctx.stage_reads(rs)
        n = ctx.snp_scan(capi.snp_params(dict(BASE_DCT), "diploid"), [(1, 80_000)])
        mat, meta, depth, count = ctx.snp_fetch()
        out.append((n, zlib.crc32(mat.tobytes()), zlib.crc32(meta.tobytes())))
        ctx.close()
    assert out[0] == out[1] == out[2] and out[0][0] > 300
