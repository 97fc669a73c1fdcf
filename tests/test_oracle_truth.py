"""CPU: the CNN restatement (oracle/cnn_oracle.py) cannot be pinned against TensorFlow here, so it is pinned against GROUND TRUTH
instead: the shipped ONT-HG002 weights, read by our own checkpoint reader and run through the restated layers on tensors of the
restated pileup code, must call the synthetic world's known SNPs with the right allele and genotype.  A wrong layer order, flatten
order, head order (A,G,T,C), kernel layout or scaling rule gives chance-level output, not 95 % recall."""
import numpy as np


def test_released_model_through_the_restatement_recovers_the_truth_snps():
    from nanocaller_b200.host import weights as W
    from nanocaller_b200.synth import make_world
    from oracle import cnn_oracle, snp_caller_oracle, snp_oracle
    w = make_world(chrom="chrT", preset="ont", contig_len=300_000, seed=31, coverage=30.0)
    tp, kinds, alts = w.truth_snps()
    truth = {int(p): (chr(a), int(k)) for p, k, a in zip(tp, kinds, alts)}
    assert len(truth) >= 300
    dct = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont", supplementary=False)
    tensors, meta = W.load_model("snp", "ONT-HG002")
    lines = []
    for ch in snp_oracle.get_chunks([("chrT", 1, 300_000, "diploid")], 1):
        pos, ref, mat, dp, freq, depth, fwd, rev = snp_oracle.get_snp_testing_candidates(w.reads, dct, ch)
        x = snp_oracle.scale_counts(mat, meta["train_coverage"], coverage=float(depth))
        ref = np.asarray(ref, np.float32)
        lines += snp_caller_oracle.diploid_records("chrT", pos, ref, cnn_oracle.snp_probs(tensors, x, ref), dp, freq, fwd, rev)
    calls = {}
    for ln in lines:
        f = ln.split("\t")
        if f[6] == "PASS":
            calls[int(f[1])] = (f[4], f[9].split(":")[0])
    hits = [p for p in calls if p in truth]
    recall, precision = len(hits) / len(truth), len(hits) / len(calls)
    right = sum(1 for p in hits if calls[p][0] == truth[p][0] and calls[p][1] == ("0/1" if truth[p][1] == 1 else "1/1"))
    assert recall > 0.9 and precision > 0.7, (recall, precision)          # false positives: the generator's 1 % systematic-error sites
    assert right / len(hits) > 0.95, right / len(hits)
    # control: the same weights with the tensor's candidate-base rows rotated (A,G,T,C -> G,T,C,A) lose the alleles
    rot = np.asarray(mat)[:, [0, 2, 3, 4, 1]]
    probs_rot = cnn_oracle.snp_probs(tensors, snp_oracle.scale_counts(rot, meta["train_coverage"], coverage=float(depth)), ref)
    lines_rot = snp_caller_oracle.diploid_records("chrT", pos, ref, probs_rot, dp, freq, fwd, rev)
    right_rot = 0
    for ln in lines_rot:
        f = ln.split("\t")
        p = int(f[1])
        if f[6] == "PASS" and p in truth and f[4] == truth[p][0]:
            right_rot += 1
    assert right_rot < 0.5 * right


def _truth_indels(w):
    idx = np.nonzero(w.indel & 63)[0]
    return {int(p) + 1: (int(w.indel[p] & 63), bool(w.indel[p] & 0x40), bool(w.indel[p] & 0x80)) for p in idx}


def score_indel_calls(lines, truth):
    """-> (truth indels found with the exact length within the candidate window, of those with the right zygosity)."""
    found = right = 0
    used = set()
    for ln in lines:
        f = ln.split("\t")
        p, ref, alts, gt = int(f[1]), f[3], f[4].split(","), f[9].split(":")[0]
        for tpos, (L, ins, hom) in truth.items():
            if tpos not in used and abs(tpos - p) <= 45 and (L if ins else -L) in [len(a) - len(ref) for a in alts]:
                found += 1
                used.add(tpos)
                right += (gt == "1/1") == hom
                break
    return found, right


def test_indel_restatement_with_the_own_msa_recovers_the_truth_indels():
    """MUSCLE and parasail are replaced by this repository's own alignments (unpinnable against the reference): what can be checked
    is that the whole indel path — scan, slices, star alignment tensors, released ONT-HG002 indel CNN, allele extraction, genotype
    rule — finds the synthetic world's indels with the exact length and zygosity."""
    from nanocaller_b200.host import weights as W
    from nanocaller_b200.synth import make_world
    from oracle import cnn_oracle, indel_caller_oracle, indel_oracle, snp_oracle
    w = make_world(chrom="chrT", preset="ont", contig_len=80_000, seed=33, coverage=30.0, indel_every=1500, indel_maxlen=12)
    truth = _truth_indels(w)
    assert len(truth) >= 40
    idct = dict(mincov=4, maxcov=160, seq="ont", del_t=0.6, ins_t=0.4, impute_indel_phase=False, supplementary=False, win_size=40, small_win_size=4)
    it, _ = W.load_model("indel", "ONT-HG002")
    lines = []
    for ch in snp_oracle.get_chunks([("chrT", 1, 80_000, "diploid")], 1, 100_000):
        pos, x0, x1, x2, alleles, phase = indel_oracle.get_indel_testing_candidates(w.reads, idct, ch)
        if len(pos):
            lines += indel_caller_oracle.diploid_records("chrT", pos, cnn_oracle.indel_model(it, np.hstack([x0, x1, x2]).astype(np.float32)), alleles, phase)
    found, right = score_indel_calls(lines, truth)
    assert found / len(truth) > 0.8 and right / found > 0.9 and len(lines) < 1.5 * len(truth), (found, right, len(lines), len(truth))
