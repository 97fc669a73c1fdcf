"""GPU end to end: BAM-native arrays -> K0/K1/K2 -> fused scaling + CNN -> host genotype/record code, against the
oracle pipeline (oracle tensors -> scale_counts -> fp32 CNN -> restated snpCaller record code).  Records must be
identical except for QUAL / PR digits that move with the <= 1e-4 probability tolerance (SURVEY D4 comparator)."""
import numpy as np
import pytest

from tests.golden_util import golden_chunk, load_case

pytestmark = pytest.mark.gpu


def _oracle_lines(name, model, haploid):
    from nanocaller_b200.host import weights as W
    from oracle import cnn_oracle, snp_caller_oracle, snp_oracle
    rs, dct, chunks, bed, g = load_case(name)
    tensors, meta = W.load_model("snp", model)
    tc = 30.0 if haploid else meta["train_coverage"]
    lines = []
    for ci, ch in enumerate(chunks):
        w = golden_chunk(g, ci)
        if len(w["pos"]) == 0:
            continue
        x = snp_oracle.scale_counts(w["mat"], tc, coverage=float(w["depth"]))
        ref = w["ref"].astype(np.float32)
        if haploid:
            probs = cnn_oracle.haploid_snp_model(tensors, x, ref)
            lines += snp_caller_oracle.haploid_records(ch["chrom"], w["pos"], w["ref"], probs, w["dp"], w["freq"])
        else:
            probs = cnn_oracle.snp_probs(tensors, x, ref)
            lines += snp_caller_oracle.diploid_records(ch["chrom"], w["pos"], w["ref"], probs, w["dp"], w["freq"], w["fwd"], w["rev"])
    return lines


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("name,model,haploid", [("ont_diploid", "ONT-HG002", False), ("hifi_pacbio", "CCS-HG002", False),
                                                ("haploid", "haploid", True)])
def test_vcf_records_match_oracle_pipeline(name, model, haploid, impl):
    from nanocaller_b200.host import snp_caller, sources, weights as W
    from nanocaller_b200.host.vcf_compare import compare_records
    rs, dct, chunks, bed, g = load_case(name)
    sources.unregister_all()
    sources.register_source("mem://bam", rs)
    params = dict(dct, sam_path="mem://bam", fasta_path="mem://bam", disable_coverage_normalization=False)
    tensors, meta = W.load_model("snp", model)
    got = snp_caller.call_chunks(params, chunks, (tensors, meta["train_coverage"]), hap_weights=tensors if haploid else None, impl=impl)
    want = _oracle_lines(name, model, haploid)
    res = compare_records(got, want, tol=1e-4)
    assert not res["mismatch"], res["mismatch"][:3]
    assert res["identical"] + res["numeric_only"] + res["borderline"] == len(want) > 0
    assert res["borderline"] <= max(1, len(want) // 500)
    assert res["identical"] >= (0.5 if haploid else 0.9) * len(want)      # haploid QUAL = -100 log10(1-p) magnifies 1e-6 shifts


def test_gpu_path_recovers_the_truth_snps():
    """Ground truth, not only the oracle: the GPU path (pileup kernels + tensor-core CNN + library record code) with the released
    ONT-HG002 weights calls the synthetic world's known SNPs with the right allele and genotype."""
    from nanocaller_b200.host import snp_caller, snp_pileups, sources, weights as W
    from nanocaller_b200.synth import make_world
    w = make_world(chrom="chrT", preset="ont", contig_len=300_000, seed=31, coverage=30.0)
    tp, kinds, alts = w.truth_snps()
    truth = {int(p): (chr(a), int(k)) for p, k, a in zip(tp, kinds, alts)}
    sources.unregister_all()
    snp_pileups._staged.clear()
    sources.register_source("mem://truth", w.reads)
    params = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont", supplementary=False,
                  sam_path="mem://truth", fasta_path="mem://truth")
    tensors, meta = W.load_model("snp", "ONT-HG002")
    lines = snp_caller.call_chunks(params, [{"chrom": "chrT", "start": 1, "end": 300_000, "ploidy": "diploid"}], (tensors, meta["train_coverage"]))
    calls = {int(f[1]): (f[4], f[9].split(":")[0]) for f in (ln.split("\t") for ln in lines) if f[6] == "PASS"}
    hits = [p for p in calls if p in truth]
    right = sum(1 for p in hits if calls[p][0] == truth[p][0] and calls[p][1] == ("0/1" if truth[p][1] == 1 else "1/1"))
    assert len(hits) / len(truth) > 0.9 and len(hits) / len(calls) > 0.7 and right / len(hits) > 0.95
