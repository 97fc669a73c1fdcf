"""GPU parity of the indel feature path through the C-ABI against fixtures produced by the UNMODIFIED reference
generate_indel_pileups.py (run over the stand-in muscle / parasail, see tests/golden/make_golden_indel.py):
candidate positions, the three float tensors bit for bit, allele strings and phase sets; then the indel CNN on them."""
import json

import numpy as np
import pytest

from tests.test_indel_oracle_golden import CASES, load_indel_case

pytestmark = pytest.mark.gpu


def _check(got, g, ci, tag):
    pos, x0, x1, x2, alleles, phase = got
    want_pos = g["c%d_pos" % ci]
    assert list(pos) == list(want_pos), (tag, list(pos)[:10], list(want_pos)[:10])
    if len(want_pos) == 0:
        return
    for k, x in (("x0", x0), ("x1", x1), ("x2", x2)):
        w = g["c%d_%s" % (ci, k)].astype(np.float64)
        bad = np.nonzero((np.asarray(x) != w).reshape(len(w), -1).any(1))[0]
        assert len(bad) == 0, "%s %s: %d/%d tensors differ, first site %d" % (tag, k, len(bad), len(w), bad[0])
    assert json.loads(json.dumps(alleles)) == json.loads(str(g["c%d_alleles" % ci])), tag
    assert list(phase) == json.loads(str(g["c%d_phase" % ci])), tag


@pytest.mark.parametrize("name", CASES)
def test_drop_in_matches_reference_golden(name):
    from nanocaller_b200.host import indel_pileups, sources
    rs, dct, chunks, g = load_indel_case(name)
    sources.unregister_all()
    sources.register_source("mem://bam", rs)
    d = dict(dct, fasta_path="mem://bam")
    for ci, chunk in enumerate(chunks):
        if chunk["ploidy"] == "haploid":
            pos, x, alleles = indel_pileups.get_indel_testing_candidates_haploid(d, dict(chunk, sam_path="mem://bam"))
            got = (pos, x, x, x, alleles, [])
        else:
            got = indel_pileups.get_indel_testing_candidates(d, dict(chunk, sam_path="mem://bam"))
        _check(got, g, ci, (name, ci))


@pytest.mark.parametrize("name", CASES)
def test_batched_chunks_match_reference_golden(name):
    from nanocaller_b200.host import indel_pileups, snp_pileups
    rs, dct, chunks, g = load_indel_case(name)
    hap = chunks[0]["ploidy"] == "haploid"
    res = indel_pileups.candidates_for_chunks(snp_pileups.context(0), rs, dct, chunks, haploid=hap)
    for ci in range(len(chunks)):
        got = res[ci] if not hap else (res[ci][0], res[ci][1], res[ci][1], res[ci][1], res[ci][2], [])
        _check(got, g, ci, (name, ci, "batched"))


def test_indel_cnn_on_device_tensors():
    """hstack of the three tensors -> Indel_model (indelCaller.py:83-85) vs the fp32 oracle."""
    from nanocaller_b200.host import indel_pileups, snp_pileups, weights as W
    from oracle import cnn_oracle
    rs, dct, chunks, g = load_indel_case("indel_ont")
    ctx = snp_pileups.context(0)
    pos, x0, x1, x2, alleles, phase = indel_pileups.candidates_for_chunks(ctx, rs, dct, chunks)[0]
    x = np.hstack([x0, x1, x2]).astype(np.float32)
    tensors, _ = W.load_model("indel", "ONT-HG002")
    ctx.load_indel_weights(W.pack_indel_blob(tensors), False)
    got = ctx.indel_model_forward(x, haploid=False, impl=1)
    want = cnn_oracle.indel_model(tensors, x)
    assert np.abs(got - want).max() < 1e-4


def test_call_chunk_records_match_oracle_pipeline():
    """indel_run equivalent for one chunk: GPU scan/build/CNN + host records vs oracle tensors -> fp32 CNN -> restated records."""
    from nanocaller_b200.host import indel_caller, sources, weights as W
    from oracle import cnn_oracle, indel_caller_oracle, indel_oracle
    rs, dct, chunks, g = load_indel_case("indel_ont")
    sources.unregister_all()
    sources.register_source("mem://bam", rs)
    tensors, _ = W.load_model("indel", "ONT-HG002")
    chunk = dict(chunks[0], sam_path="mem://bam")
    got = indel_caller.call_chunk(dict(dct, fasta_path="mem://bam"), chunk, tensors)
    pos, x0, x1, x2, alleles, phase = indel_oracle.get_indel_testing_candidates(rs, dct, chunks[0])
    probs = cnn_oracle.indel_model(tensors, np.hstack([x0, x1, x2]).astype(np.float32))
    want = indel_caller_oracle.diploid_records(chunks[0]["chrom"], pos, probs, alleles, phase)
    assert len(got) == len(want) > 0
    for a, b in zip(got, want):
        fa, fb = a.split("\t"), b.split("\t")
        assert fa[:5] == fb[:5] and fa[6:9] == fb[6:9], (a, b)
        assert abs(float(fa[5]) - float(fb[5])) < 0.02
        assert fa[9].split(":")[0] == fb[9].split(":")[0]


def test_gpu_indel_path_recovers_the_truth_indels():
    """Ground truth for the product path: GPU scan / slices / star alignment / indel CNN + host allele and genotype code on a synthetic
    contig with known indels (exact length, zygosity)."""
    from nanocaller_b200.host import indel_caller, snp_pileups, sources, weights as W
    from nanocaller_b200.synth import make_world
    from tests.test_oracle_truth import _truth_indels, score_indel_calls
    w = make_world(chrom="chrT", preset="ont", contig_len=400_000, seed=33, coverage=30.0, indel_every=1500, indel_maxlen=12)
    truth = _truth_indels(w)
    sources.unregister_all()
    snp_pileups._staged.clear()
    sources.register_source("mem://truth", w.reads)
    idct = dict(mincov=4, maxcov=160, seq="ont", del_t=0.6, ins_t=0.4, impute_indel_phase=False, supplementary=False, win_size=40, small_win_size=4,
                fasta_path="mem://truth")
    it, _ = W.load_model("indel", "ONT-HG002")
    chunks = [{"chrom": "chrT", "start": s, "end": min(400_000, s + 100_000), "ploidy": "diploid", "sam_path": "mem://truth"} for s in range(1, 400_000, 100_000)]
    lines = indel_caller.call_chunks(idct, chunks, it)
    found, right = score_indel_calls(lines, truth)
    assert len(truth) > 200 and found / len(truth) > 0.85 and right / found > 0.95 and len(lines) < 1.4 * len(truth), (found, right, len(lines), len(truth))


@pytest.mark.parametrize("impl", [0, 1])
def test_device_resident_indel_forward_matches_oracle(impl):
    """nc_indel_forward: the CNN on the tensors where nc_indel_build left them (no host round trip) vs the fp32 oracle on the fetched tensors."""
    from nanocaller_b200.host import indel_pileups, snp_pileups, weights as W
    from oracle import cnn_oracle
    rs, dct, chunks, g = load_indel_case("indel_ont")
    ctx = snp_pileups.context(0)
    meta, tensors, cns = indel_pileups.scan_build(ctx, rs, dct, chunks, want_tensors=True)
    w, _ = W.load_model("indel", "ONT-HG002")
    ctx.load_indel_weights(W.pack_indel_blob(w), False)
    got = ctx.indel_forward(impl=impl)
    keep = indel_pileups.kept_sites(meta, False)
    want = cnn_oracle.indel_model(w, tensors.reshape(len(meta), 15, 128, 2))
    assert keep.sum() > 10 and got.shape == (len(meta), 4)
    assert float(np.abs(got[keep] - np.asarray(want)[keep]).max()) < 1e-4
    assert float(np.abs(got - np.asarray(want)).max()) < 1e-4           # sites msa() rejected carry zero tensors: still a defined input


@pytest.mark.parametrize("preset,seq", [("ont", "ont"), ("hifi", "pacbio")])
def test_device_allele_prediction_equals_host_alignment(preset, seq):
    """indel_allele_kernel (affine NW + traceback + the reference's CIGAR walk on the GPU) against nc_allele_predict_batch (the same on
    host threads) for every kept (site, group) of a synthetic contig: identical allele lengths, none deferred to the host."""
    from nanocaller_b200.host import indel_pileups, snp_pileups
    from nanocaller_b200.synth import make_world
    rs = make_world(chrom="chrQ", preset=preset, contig_len=300_000, seed=41, coverage=30.0, indel_every=1200, indel_maxlen=40).reads
    dct = dict(mincov=4, maxcov=160, seq=seq, del_t=0.6, ins_t=0.4, impute_indel_phase=False, supplementary=False, win_size=40, small_win_size=4)
    chunks = [{"chrom": "chrQ", "start": s, "end": min(300_000, s + 100_000), "ploidy": "diploid"} for s in range(1, 300_000, 100_000)]
    ctx = snp_pileups.context(0)
    snp_pileups._staged.clear()
    meta, _, cns = indel_pileups.scan_build(ctx, rs, dct, chunks, want_tensors=False)
    dev = ctx.indel_fetch_alleles()
    host = indel_pileups.AllelePredictions(rs, dct, meta, cns, False)
    devp = indel_pileups.AllelePredictions(rs, dct, meta, cns, False, device_lengths=dev)
    assert len(host.site) > 300 and not (dev == -2).any()
    assert np.array_equal(host.ref_out, devp.ref_out) and np.array_equal(host.alt_out, devp.alt_out) and host.strings() == devp.strings()
    assert (host.ref_out >= 0).sum() > 50 and (host.ref_out < 0).sum() > 50
    kept = indel_pileups.kept_sites(meta, False)
    assert (dev[~kept] == -1).all()


@pytest.mark.parametrize("preset,seq", [("ont", "ont"), ("hifi", "pacbio")])
def test_paired_alignment_kernel_equals_one_slice_per_warp_kernel(preset, seq):
    """indel_align2_kernel (two slices of a site per warp in 16-bit halves, tie bits as directions; six-column strips for the
    161-column window, nine-column strips for the HiFi preset's 261) against indel_align_kernel (one slice per warp, explicit
    compares; the kernel the reference fixtures pinned first) on every site of a synthetic contig with indels up to 40 bases, reads
    ending inside windows and the last window cut by the contig end: identical tensors, consensus strings and site records."""
    import os
    from nanocaller_b200.host import indel_pileups, snp_pileups
    from nanocaller_b200.synth import make_world
    rs = make_world(chrom="chrP", preset=preset, contig_len=400_000, seed=43, coverage=25.0, indel_every=700, indel_maxlen=40).reads
    dct = dict(mincov=4, maxcov=160, seq=seq, del_t=0.6, ins_t=0.4, impute_indel_phase=False, supplementary=False, win_size=40, small_win_size=4)
    chunks = [{"chrom": "chrP", "start": s, "end": min(400_000, s + 100_000), "ploidy": "diploid"} for s in range(1, 400_000, 100_000)]
    ctx = snp_pileups.context(0)
    out = {}
    try:
        for mode in ("0", "1"):
            os.environ["NC_INDEL_ALIGN_SCALAR"] = mode
            snp_pileups._staged.clear()
            meta, tensors, cns = indel_pileups.scan_build(ctx, rs, dct, chunks)
            out[mode] = (meta.copy(), np.array(tensors, copy=True), np.array(cns, copy=True), ctx.indel_fetch_alleles().copy())
    finally:
        os.environ.pop("NC_INDEL_ALIGN_SCALAR", None)
    a, b = out["0"], out["1"]
    assert len(a[0]) > 400 and indel_pileups.kept_sites(a[0], False).sum() > 300
    assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
