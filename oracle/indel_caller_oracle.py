"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restatement of the indel genotype decision and VCF record formatting of the reference worker
(indelCaller.py:74-152 diploid, :159-182 haploid), taking CNN probabilities and allele strings as input.
Kept line for line (including the `prev` overlap suppression and the QUAL quirk: the min(99, ...) value is discarded,
SURVEY appendix F.12).  Float semantics as in the reference's environment: `-10*np.log10(1e-6 + p)` on float32
probabilities is evaluated in float64 for the scalar expressions and in float32 for `qual_all` (unused).
Pinned: tests/test_indel_records_golden.py holds it to the records the UNMODIFIED `indelCaller.indel_run` wrote over oracle/shim
(tests/golden/records_indel_*.vcf.txt, made by tests/golden/make_golden_indel_records.py)."""
import numpy as np


def diploid_records(chrom, pos, probs, alleles_seq, phase):
    batch_prob_all = np.asarray(probs, np.float32)
    batch_pred_all = np.argmax(batch_prob_all, axis=1)
    out = []
    prev = 0
    for j in range(len(batch_pred_all)):
        if pos[j] > prev:
            if batch_prob_all[j, 0] <= 0.95:
                q = -10 * np.log10(1e-6 + np.float64(batch_prob_all[j, 0]))
                allele0_data, allele1_data, allele_total_data = alleles_seq[j]
                if batch_pred_all[j] == 1 and allele_total_data[0]:
                    gq = -10 * np.log10(1 + 1e-6 - np.float64(batch_prob_all[j, 1]))
                    out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ\t1/1:%.2f\n' % (chrom, pos[j], allele_total_data[0], allele_total_data[1], q, gq))
                    prev = pos[j] + max(len(allele_total_data[0]), len(allele_total_data[1]))
                else:
                    if allele0_data[0] and allele1_data[0]:
                        if allele0_data[0] == allele1_data[0] and allele0_data[1] == allele1_data[1]:
                            gq = -10 * np.log10(1 + 1e-6 - np.float64(batch_prob_all[j, 1]))
                            out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ\t1/1:%.2f\n' % (chrom, pos[j], allele0_data[0], allele0_data[1], q, gq))
                            prev = pos[j] + max(len(allele0_data[0]), len(allele0_data[1]))
                        else:
                            ref1, alt1 = allele0_data
                            ref2, alt2 = allele1_data
                            l = min(len(ref1), len(ref2))
                            if len(ref1) > len(ref2):
                                ref = ref1
                                alt2 = alt2 + ref1[l:]
                            else:
                                ref = ref2
                                alt1 = alt1 + ref2[l:]
                            gq = -10 * np.log10(1 + 1e-6 - np.float64(batch_prob_all[j, 3]))
                            if phase[j]:
                                out.append('%s\t%d\t.\t%s\t%s,%s\t%.2f\tPASS\t.\tGT:GQ:PS\t1|2:%.2f:%d\n' % (chrom, pos[j], ref, alt1, alt2, q, gq, phase[j]))
                            else:
                                out.append('%s\t%d\t.\t%s\t%s,%s\t%.2f\tPASS\t.\tGT:GQ\t1|2:%.2f\n' % (chrom, pos[j], ref, alt1, alt2, q, gq))
                            prev = pos[j] + max(len(ref), len(alt1), len(alt2))
                    elif allele0_data[0]:
                        gq = -10 * np.log10(1 + 1e-6 - np.float64(batch_prob_all[j, 2]))
                        if phase[j]:
                            out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ:PS\t0|1:%.2f:%d\n' % (chrom, pos[j], allele0_data[0], allele0_data[1], q, gq, phase[j]))
                        else:
                            out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ\t0|1:%.2f\n' % (chrom, pos[j], allele0_data[0], allele0_data[1], q, gq))
                        prev = pos[j] + max(len(allele0_data[0]), len(allele0_data[1]))
                    elif allele1_data[0]:
                        gq = -10 * np.log10(1 + 1e-6 - np.float64(batch_prob_all[j, 2]))
                        if phase[j]:
                            out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ:PS\t1|0:%.2f:%d\n' % (chrom, pos[j], allele1_data[0], allele1_data[1], q, gq, phase[j]))
                        else:
                            out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ\t1|0:%.2f\n' % (chrom, pos[j], allele1_data[0], allele1_data[1], q, gq))
                        prev = pos[j] + max(len(allele1_data[0]), len(allele1_data[1]))
    return out


def haploid_records(chrom, pos, probs, alleles_seq):
    batch_prob = np.asarray(probs, np.float32).reshape(-1)
    out = []
    prev = 0
    for j in range(len(batch_prob)):
        allele_total_data = alleles_seq[j]
        if pos[j] > prev and batch_prob[j] >= 0.5 and allele_total_data[0]:
            q = -100 * np.log10(1e-6 + 1 - np.float64(batch_prob[j]))
            out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ\t1/1:%.2f\n' % (chrom, pos[j], allele_total_data[0], allele_total_data[1], q, q))
            prev = pos[j] + max(len(allele_total_data[0]), len(allele_total_data[1]))
    return out
