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
