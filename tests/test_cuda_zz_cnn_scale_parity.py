"""GPU: the tensor-core SNP CNN against the fp32 ORACLE (oracle/cnn_oracle.py, not the library's own fp32 path) at scale:
every released diploid SNP model (snpCaller.py:16-34; the 16 distinct weight sets + the NanoCaller2 alias) on device-built tensors
of three synthetic contigs at 30x, 60x and 160x — more than 10^6 (site, model) evaluations — plus the haploid model at 60x.
The kernels run in their validation instantiation, which counts activations that sit on the fp16 saturation value: the hi / lo
split converts with cvt.rn.satfinite, which would clamp |x| > 65504 silently; the count must stay 0 at 160x, where the tensor
counts are largest, and for the models with the largest weights (CLR-HG002, NanoCaller1: max |w| 11)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DCT = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont", supplementary=False)
LENGTH = 2_400_000


def _saturation(ctx, haploid, on, reset=True):
    from nanocaller_b200.host import capi
    lib = capi.load_library()
    lib.nc_debug_saturation.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
    lib.nc_debug_saturation.restype = ctypes.c_int
    n = ctypes.c_int64(-1)
    rc = lib.nc_debug_saturation(ctx._h, 1 if haploid else 0, 1 if on else 0, 1 if reset else 0, ctypes.byref(n))
    assert rc == 0, lib.nc_last_error(ctx._h)
    return n.value


def _world_tensors(ctx, cov, ploidy):
    from nanocaller_b200.cli import get_chunks
    from nanocaller_b200.host import snp_pileups
    from nanocaller_b200.synth import make_world
    rs = make_world(chrom="chrP", preset="ont", contig_len=LENGTH, seed=100 + int(cov), coverage=float(cov), ploidy=1 if ploidy == "haploid" else 2).reads
    chunks = get_chunks([("chrP", 1, LENGTH, ploidy)], 1)
    snp_pileups._staged.clear()
    n = snp_pileups.scan_chunks(ctx, rs, DCT, chunks, ploidy)
    mat, meta, depth, count = ctx.snp_fetch()
    return n, mat[:, :1025].reshape(n, 5, 41, 5), meta, depth


def _oracle_probs(tensors, mat, meta, depth, tc, haploid):
    from oracle import cnn_oracle, snp_oracle
    ref = np.zeros((len(meta), 4), np.float32)
    ref[np.arange(len(meta)), meta["ref_code"]] = 1
    out = np.empty((len(meta), 4), np.float32)
    for c in np.unique(meta["chunk"]):
        sel = np.nonzero(meta["chunk"] == c)[0]
        x = snp_oracle.scale_counts(mat[sel].astype(np.int32), tc, coverage=float(depth[c]))
        for b in range(0, len(sel), 4000):
            s = sel[b:b + 4000]
            out[s] = cnn_oracle.haploid_snp_model(tensors, x[b:b + 4000], ref[s]) if haploid else cnn_oracle.snp_probs(tensors, x[b:b + 4000], ref[s])
    return out


def test_all_snp_models_match_the_fp32_oracle_at_30x_60x_160x():
    import torch
    from nanocaller_b200.host import snp_pileups, weights as W
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ctx = snp_pileups.context(0)
    names = [m for m in W.SNP_MODEL_DICT if m != "haploid"]
    assert len(names) == 17
    total, worst = 0, 0.0
    report = {}
    for cov in (30, 60, 160):
        n, mat, meta, depth = _world_tensors(ctx, cov, "diploid")
        assert n > 15_000 and int(np.abs(mat).max()) >= (25 if cov == 30 else cov // 2)
        for name in names:
            tensors, m = W.load_model("snp", name)
            ctx.load_snp_weights(W.pack_snp_blob(tensors, False), m["train_coverage"], False)
            _saturation(ctx, False, True)
            got = ctx.snp_forward(normalize=True, impl=0)
            sat = _saturation(ctx, False, False)
            want = _oracle_probs(tensors, mat, meta, depth, m["train_coverage"], False)
            err = float(np.abs(got - want).max())
            report[(name, cov)] = (err, sat)
            assert sat == 0, (name, cov, sat)
            assert err < 1e-4, (name, cov, err)
            total += n
            worst = max(worst, err)
    assert total >= 1_000_000, total
    print("\n%d (site, model) evaluations, max |dP| %.2e" % (total, worst))


def test_haploid_snp_model_matches_the_fp32_oracle_at_60x():
    from nanocaller_b200.host import snp_pileups, weights as W
    ctx = snp_pileups.context(0)
    n, mat, meta, depth = _world_tensors(ctx, 60, "haploid")
    tensors, _ = W.load_model("snp", "haploid")
    ctx.load_snp_weights(W.pack_snp_blob(tensors, True), 30.0, True)
    _saturation(ctx, True, True)
    got = ctx.snp_forward(normalize=True, impl=0)
    assert _saturation(ctx, True, False) == 0
    want = _oracle_probs(tensors, mat, meta, depth, 30.0, True)
    assert n > 5_000 and float(np.abs(got - want).max()) < 1e-4
