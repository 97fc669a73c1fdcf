"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restatement of the SNP genotype decision and VCF record formatting of the reference worker
(snpCaller.py:113-163 diploid, :183-198 haploid), taking the CNN probabilities as input.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import it.
Pinned: tests/test_snp_records_golden.py holds it to the records the UNMODIFIED `snpCaller.caller` wrote over oracle/shim
(tests/golden/records_*.vcf.txt, made by tests/golden/make_golden_records.py).

Float semantics follow the reference's pinned environment (environment.yml:9, numpy<2):
`1e-10 + 1 - probs[j,k]` mixes a Python float with a float32 scalar, which legacy promotion
evaluates in float64; QUAL is therefore min(99, -10*log10(1.0000000001 - float64(p))).
np.argsort on 4-element rows is an insertion sort, i.e. stable: among equal probabilities the
higher base index ranks higher.  (TensorFlow is unavailable here, so these rules are pinned to the
source text, not to an executed reference run.)
"""
import numpy as np

NUM_TO_BASE = {0: "A", 1: "G", 2: "T", 3: "C"}   # snpCaller.py:14


def diploid_records(chrom, pos, ref_onehot, probs, dp, freq, fwd_dp, rev_dp):
    """snpCaller.py:113-163.  probs float32 [N,4] = P(A),P(G),P(T),P(C); returns list of VCF lines."""
    batch_probs = np.asarray(probs, dtype=np.float32)
    batch_pred = np.argsort(batch_probs, axis=1, kind="stable")          # :118
    batch_ref = np.argmax(np.asarray(ref_onehot), 1)                      # :120
    batch_pred_GT = np.sum(batch_probs >= 0.5, axis=1)                    # :122
    fwd_dp = np.asarray(fwd_dp, np.float64)
    rev_dp = np.asarray(rev_dp, np.float64)
    out = []
    for j in range(len(batch_pred_GT)):
        info_field = "PR=" + ",".join("{:.4f}".format(x) for x in batch_probs[j, [0, 3, 1, 2]]) + ";FQ={:.4f}".format(freq[j])
        r = batch_ref[j]
        ref_dp = (fwd_dp[j][r], rev_dp[j][r])
        q = lambda k: min(99, -10 * np.log10(1e-10 + 1 - np.float64(batch_probs[j, k])))
        if batch_pred_GT[j] >= 2:                                         # :130
            pred1, pred2 = batch_pred[j, -1], batch_pred[j, -2]
            if pred1 == r:                                                # :132
                alt_dp = (fwd_dp[j][pred2], rev_dp[j][pred2])
                out.append("%s\t%d\t.\t%s\t%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:%.4f:%d,%d:%d,%d:%d,%d\n" % (
                    chrom, pos[j], NUM_TO_BASE[r], NUM_TO_BASE[pred2], q(pred2), "PASS", info_field, "0/1", dp[j],
                    sum(alt_dp) / dp[j], sum(ref_dp), sum(alt_dp), ref_dp[0], alt_dp[0], ref_dp[1], alt_dp[1]))
            elif pred2 == r and batch_probs[j, pred2] >= 0.5:             # :138
                alt_dp = (fwd_dp[j][pred1], rev_dp[j][pred1])
                out.append("%s\t%d\t.\t%s\t%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:%.4f:%d,%d:%d,%d:%d,%d\n" % (
                    chrom, pos[j], NUM_TO_BASE[r], NUM_TO_BASE[pred1], q(pred2), "PASS", info_field, "0/1", dp[j],
                    sum(alt_dp) / dp[j], sum(ref_dp), sum(alt_dp), ref_dp[0], alt_dp[0], ref_dp[1], alt_dp[1]))
            elif pred2 != r and pred1 != r and batch_probs[j, pred2] >= 0.5:   # :143
                alt1_dp = (fwd_dp[j][pred1], rev_dp[j][pred1])
                alt2_dp = (fwd_dp[j][pred2], rev_dp[j][pred2])
                out.append("%s\t%d\t.\t%s\t%s,%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:%.4f,%.4f:%d,%d,%d:%d,%d,%d:%d,%d,%d\n" % (
                    chrom, pos[j], NUM_TO_BASE[r], NUM_TO_BASE[pred1], NUM_TO_BASE[pred2], q(pred2), "PASS", info_field,
                    "1/2", dp[j], sum(alt1_dp) / dp[j], sum(alt2_dp) / dp[j], sum(ref_dp), sum(alt1_dp), sum(alt2_dp),
                    ref_dp[0], alt1_dp[0], alt2_dp[0], ref_dp[1], alt1_dp[1], alt2_dp[1]))
        elif batch_pred_GT[j] == 1 and r != batch_pred[j, -1] and batch_probs[j, batch_pred[j, -1]] >= 0.5:   # :150
            pred1 = batch_pred[j, -1]
            alt_dp = (fwd_dp[j][pred1], rev_dp[j][pred1])
            out.append("%s\t%d\t.\t%s\t%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:%.4f:%d,%d:%d,%d:%d,%d\n" % (
                chrom, pos[j], NUM_TO_BASE[r], NUM_TO_BASE[pred1], q(pred1), "PASS", info_field, "1/1", dp[j],
                sum(alt_dp) / dp[j], sum(ref_dp), sum(alt_dp), ref_dp[0], alt_dp[0], ref_dp[1], alt_dp[1]))
        else:
            if batch_pred_GT[j] == 1 and r == batch_pred[j, -1]:          # :157
                pred1 = batch_pred[j, -1]
                out.append("%s\t%d\t.\t%s\t%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:.:.:.:.\n" % (
                    chrom, pos[j], NUM_TO_BASE[r], ".", q(pred1), "REF", info_field, "./.", dp[j]))
            else:                                                         # :161
                out.append("%s\t%d\t.\t%s\t%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:.:.:.:.\n" % (
                    chrom, pos[j], NUM_TO_BASE[r], ".", 0, "LOW", info_field, "./.", dp[j]))
    return out


def haploid_records(chrom, pos, ref_onehot, probs, dp, freq):
    """snpCaller.py:183-198.  probs float32 [N,4] softmax over A,G,T,C."""
    batch_probs = np.asarray(probs, dtype=np.float32)
    batch_ref = np.argmax(np.asarray(ref_onehot), 1)
    batch_pred = np.argmax(batch_probs, 1)
    out = []
    for j in range(len(batch_pred)):
        pred = batch_pred[j]
        info_field = "PR=" + ",".join("{:.4f}".format(x) for x in batch_probs[j, [0, 3, 1, 2]]) + ";FQ={:.4f}".format(freq[j])
        qual = min(999, -100 * np.log10(1e-10 + 1 - np.float64(batch_probs[j, pred])))
        out.append("%s\t%d\t.\t%s\t%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:%.4f:.:.:.\n" % (
            chrom, pos[j], NUM_TO_BASE[batch_ref[j]], NUM_TO_BASE[pred], qual,
            "PASS" if pred != batch_ref[j] else "REF", info_field, "1/1", dp[j], freq[j]))
    return out
