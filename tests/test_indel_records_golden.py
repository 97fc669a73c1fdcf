"""Pins the indel decision + record restatements against the UNMODIFIED reference worker `indelCaller.indel_run`
(indelCaller.py:41-189) run over oracle/shim with the released weights (tests/golden/make_golden_indel_records.py): the models
are the reference's own model_architect_indel*.py classes (only conv / dense / selu / softmax / sigmoid come from the stand-in
tensorflow), hstack, batching, genotype rules, `prev` suppression and record text are its code.  CPU only.
Checked: the oracle chain (golden tensors -> cnn_oracle -> indel_caller_oracle) and the host's record code
(host/indel_caller.records_from_calls / haploid_records_from_calls)."""
import json
import os

import numpy as np
import pytest

from nanocaller_b200.host import indel_caller, weights as W
from oracle import cnn_oracle, indel_caller_oracle
from tests.test_indel_oracle_golden import load_indel_case

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"indel_ont": "ONT-HG002", "indel_haploid": "ONT-HG002", "indel_impute_hifi": "CCS-HG002", "indel_sub": "ONT-HG002",
         "indel_impute_ont": "ONT-HG002"}


def _same(got, want, tag):
    assert len(got) == len(want) > 20, (tag, len(got), len(want))
    exact = 0
    for a, b in zip(got, want):
        fa, fb = a.rstrip("\n").split("\t"), b.rstrip("\n").split("\t")
        assert fa[:5] == fb[:5] and fa[6:9] == fb[6:9], (tag, a, b)
        sa, sb = fa[9].split(":"), fb[9].split(":")
        assert sa[0] == sb[0] and sa[2:] == sb[2:], (tag, a, b)                  # genotype and phase set
        assert abs(float(fa[5]) - float(fb[5])) <= 0.011 and abs(float(sa[1]) - float(sb[1])) <= 0.011, (tag, a, b)   # QUAL / GQ: last printed digit
        exact += a == b
    assert exact >= 0.9 * len(want), (tag, exact, len(want))


@pytest.mark.parametrize("name", sorted(CASES))
def test_indel_records_match_the_unmodified_reference_worker(name):
    rs, dct, chunks, g = load_indel_case(name)
    lines = open(os.path.join(GOLDEN_DIR, "records_%s.vcf.txt" % name)).read().split("\n")
    assert lines[0] == "# " + rs.checksum()
    want = [ln + "\n" for ln in lines[1:] if ln]
    tensors, _ = W.load_model("indel", CASES[name])
    hap, _ = W.load_model("indel", "haploid")
    got_o, got_h = [], []
    for ci, ch in enumerate(chunks):
        pos = g["c%d_pos" % ci]
        if len(pos) == 0:
            continue
        alleles = json.loads(str(g["c%d_alleles" % ci]))
        if ch["ploidy"] == "haploid":
            probs = cnn_oracle.haploid_indel_model(hap, g["c%d_x2" % ci].astype(np.float32))
            got_o += indel_caller_oracle.haploid_records(ch["chrom"], pos, probs, alleles)
            got_h += indel_caller.haploid_records_from_calls(ch["chrom"], pos, probs, alleles)
        else:
            phase = json.loads(str(g["c%d_phase" % ci]))
            x = np.hstack([g["c%d_x0" % ci], g["c%d_x1" % ci], g["c%d_x2" % ci]]).astype(np.float32)
            probs = cnn_oracle.indel_model(tensors, x)
            got_o += indel_caller_oracle.diploid_records(ch["chrom"], pos, probs, alleles, phase)
            got_h += indel_caller.records_from_calls(ch["chrom"], pos, probs, alleles, phase)
    _same(got_o, want, (name, "oracle"))
    _same(got_h, want, (name, "host"))
