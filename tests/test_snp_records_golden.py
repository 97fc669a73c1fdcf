"""Pins the SNP decision + record restatements against the UNMODIFIED reference worker.

tests/golden/records_<case>.vcf.txt hold what `nanocaller_src.snpCaller.caller` (snpCaller.py:57-198) wrote when it was run
unchanged over oracle/shim (pysam, intervaltree and the stand-in tensorflow: the models are the reference's own
model_architect*.py classes, only conv / dense / selu / softmax come from the shim) with the released weights
(tests/golden/make_golden_records.py).  Checked here, CPU only:
  * oracle chain  snp_oracle tensors -> scale_counts -> cnn_oracle -> snp_caller_oracle records   (what the GPU tests hold the
    product to), and
  * the host's own record code (host/snp_caller.records_from_calls) and the library's threaded formatter the product path uses
    (nc_format_snp_records) on the same probabilities.
The fixtures were made under NumPy 2 (the coverage scaling multiplies in float64 there, in float32 under the reference's pinned
NumPy < 2, snpCaller.py:96), so QUAL / PR digits may differ in the last place: `compare_records` allows that and nothing else."""
import os

import numpy as np
import pytest

from nanocaller_b200.host import weights as W
from nanocaller_b200.host.vcf_compare import compare_records
from oracle import cnn_oracle, snp_caller_oracle, snp_oracle
from tests.golden_util import load_case

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"ont_diploid": {}, "haploid": {}, "ont_subregion_bed": {"disable_coverage_normalization": True}, "lowcov": {},
         "hifi_pacbio": {"snp_model": "CCS-HG002"}, "haploid_nonorm": {"_case": "haploid", "disable_coverage_normalization": True}}
_models = {}


def _model(name):
    if name not in _models:
        _models[name] = W.load_model("snp", name)
    return _models[name]


def _fixture(name, rs):
    lines = open(os.path.join(GOLDEN_DIR, "records_%s.vcf.txt" % name)).read().split("\n")
    assert lines[0] == "# " + rs.checksum()
    return [ln + "\n" for ln in lines[1:] if ln]


def _chain(name, over, record_fns):
    rs, dct, chunks, bed, g = load_case(over.get("_case", name))
    tensors, meta = _model(over.get("snp_model", "ONT-HG002"))
    hap, _ = _model("haploid")
    out = [[] for _ in record_fns]
    for ch in chunks:
        pos, ref, mat, dp, freq, depth, fwd, rev = snp_oracle.get_snp_testing_candidates(rs, dct, ch, bed_intervals=bed.get(ch["chrom"]) if bed else None)
        if len(pos) == 0:
            continue
        ref = np.asarray(ref, np.float32)
        haploid = ch["ploidy"] == "haploid"
        tc = 30.0 if haploid else meta["train_coverage"]
        if over.get("disable_coverage_normalization"):
            x = snp_oracle.scale_counts(mat, tc, dp=dp)
        else:
            x = snp_oracle.scale_counts(mat, tc, coverage=float(depth))
        probs = cnn_oracle.haploid_snp_model(hap, x, ref) if haploid else cnn_oracle.snp_probs(tensors, x, ref)
        for k, fn in enumerate(record_fns):
            out[k] += fn(ch["chrom"], haploid, pos, ref, probs, dp, freq, fwd, rev)
    return rs, out


def _oracle_records(chrom, haploid, pos, ref, probs, dp, freq, fwd, rev):
    if haploid:
        return snp_caller_oracle.haploid_records(chrom, pos, ref, probs, dp, freq)
    return snp_caller_oracle.diploid_records(chrom, pos, ref, probs, dp, freq, fwd, rev)


def _host_records(chrom, haploid, pos, ref, probs, dp, freq, fwd, rev):
    from nanocaller_b200.host import snp_caller
    return snp_caller.records_from_calls(chrom, pos, np.argmax(ref, 1), probs, dp, freq, fwd, rev, ploidy="haploid" if haploid else "diploid")


def _library_records(chrom, haploid, pos, ref, probs, dp, freq, fwd, rev):
    """nc_format_snp_records: the formatter the product path uses (host threads inside libnanocaller_b200.so)."""
    from nanocaller_b200.host import capi
    alt = np.rint(np.asarray(freq, np.float64) * np.asarray(dp)).astype(np.int32)          # freq = alt / dp (generate_SNP_pileups.py:166)
    blob, off, _ = capi.format_snp_records(chrom, pos, np.argmax(ref, 1), probs, dp, alt, fwd, rev, haploid=haploid, threads=2)
    return [blob[off[i]:off[i + 1]].decode() for i in range(len(off) - 1)]


@pytest.mark.parametrize("name", sorted(CASES))
def test_records_match_the_unmodified_reference_worker(name):
    rs, (got_o, got_h, got_l) = _chain(name, CASES[name], [_oracle_records, _host_records, _library_records])
    want = _fixture(name, rs)
    assert len(want) > 300
    for tag, got in (("oracle", got_o), ("host", got_h), ("library", got_l)):
        res = compare_records(got, want, tol=2e-6)
        assert not res["mismatch"], (tag, res["mismatch"][:2])
        assert res["borderline"] == 0 and res["identical"] >= 0.94 * len(want), (tag, res["identical"], res["numeric_only"], len(want))
