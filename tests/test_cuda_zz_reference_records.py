"""GPU product path against the records written by the UNMODIFIED reference worker (tests/golden/records_*.vcf.txt, made by
running `snpCaller.caller` over the I/O and TensorFlow stand-ins with the released weights, tests/golden/make_golden_records.py):
BAM-native arrays -> K0/K1/K2 -> fused scaling + tcgen05 CNN -> record text, compared with the SURVEY D4 comparator.
(The CPU suite holds the oracle restatements to the same fixtures; this closes the chain without the oracle in between.)"""
import os

import numpy as np
import pytest

from tests.golden_util import load_case

pytestmark = pytest.mark.gpu

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,model,haploid,disable", [("ont_diploid", "ONT-HG002", False, False), ("haploid", "haploid", True, False),
                                                        ("haploid_nonorm", "haploid", True, True)])
def test_product_records_match_the_reference_worker(name, model, haploid, disable):
    from nanocaller_b200.host import snp_caller, sources, weights as W
    from nanocaller_b200.host.vcf_compare import compare_records
    rs, dct, chunks, bed, g = load_case("haploid" if name == "haploid_nonorm" else name)
    lines = open(os.path.join(GOLDEN_DIR, "records_%s.vcf.txt" % name)).read().split("\n")
    assert lines[0] == "# " + rs.checksum()
    want = [ln + "\n" for ln in lines[1:] if ln]
    sources.unregister_all()
    sources.register_source("mem://bam", rs)
    params = dict(dct, sam_path="mem://bam", fasta_path="mem://bam", disable_coverage_normalization=disable)
    if bed is not None:
        sources.register_bed("mem://bed", bed)
        params["exclude_bed"] = "mem://bed"
    tensors, meta = W.load_model("snp", model)
    got = snp_caller.call_chunks(params, chunks, (tensors, meta["train_coverage"]), hap_weights=tensors if haploid else None)
    res = compare_records(got, want, tol=1e-4)
    n = len(want)
    assert n > 300 and len(got) == n
    assert len(res["mismatch"]) + res["borderline"] <= max(2, n // 300), (res["mismatch"][:3], res["borderline"])
    assert res["identical"] >= (0.3 if haploid else 0.75) * n          # the rest: QUAL / PR digits within the 1e-4 probability tolerance
