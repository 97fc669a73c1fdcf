"""Host mirror of the per-chunk body of snpCaller.caller (snpCaller.py:86-198): candidate tensors -> coverage
scaling -> CNN -> genotype decision -> VCF record lines.

Two ways in:
  * `call_chunks(...)`  the product path — one scan + one fused forward for all chunks of a contig on the GPU,
    then record formatting on the host from the 56-byte call records;
  * `records_from_calls(...)`  the decision/formatting step alone (numpy, vectorised), used by both.

Float semantics follow the reference's pinned environment (environment.yml:9, numpy<2): QUAL is computed in
float64 from float32 probabilities; `np.argsort` on 4-element rows is stable."""
import numpy as np

from . import capi, snp_pileups, sources

NUM_TO_BASE = np.array(["A", "G", "T", "C"])      # snpCaller.py:14


def _qual(p, cap, mult):
    with np.errstate(divide="ignore"):
        return np.minimum(cap, -mult * np.log10(1e-10 + 1 - p.astype(np.float64)))


def records_from_calls(chrom, pos, ref_code, probs, dp, freq, fwd_dp, rev_dp, ploidy="diploid"):
    """-> list of VCF lines, one per candidate, in input order (snpCaller.py:113-163 / :183-198)."""
    n = len(pos)
    if n == 0:
        return []
    probs = np.asarray(probs, np.float32)
    ref = np.asarray(ref_code, np.int64)
    pr = probs[:, [0, 3, 1, 2]]                                         # PR= in order A,C,G,T (:127)
    info = ["PR=%.4f,%.4f,%.4f,%.4f;FQ=%.4f" % (a, b, c, d, f) for (a, b, c, d), f in zip(pr.tolist(), np.asarray(freq, np.float64).tolist())]
    refb = NUM_TO_BASE[ref]
    out = []
    if ploidy == "haploid":
        pred = np.argmax(probs, 1)
        q = _qual(probs[np.arange(n), pred], 999, 100)
        flt = np.where(pred != ref, "PASS", "REF")
        for j in range(n):
            out.append("%s\t%d\t.\t%s\t%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:%.4f:.:.:.\n" % (
                chrom, pos[j], refb[j], NUM_TO_BASE[pred[j]], q[j], flt[j], info[j], "1/1", dp[j], freq[j]))
        return out
    order = np.argsort(probs, axis=1, kind="stable")
    p1, p2 = order[:, -1], order[:, -2]
    k = (probs >= 0.5).sum(1)
    fwd = np.asarray(fwd_dp, np.float64)
    rev = np.asarray(rev_dp, np.float64)
    ar = np.arange(n)
    q1 = _qual(probs[ar, p1], 99, 10)
    q2 = _qual(probs[ar, p2], 99, 10)
    prob2 = probs[ar, p2]
    for j in range(n):
        r = ref[j]
        rf, rr = fwd[j, r], rev[j, r]
        head = "%s\t%d\t.\t%s\t" % (chrom, pos[j], refb[j])
        if k[j] >= 2:
            a1, a2 = p1[j], p2[j]
            if a1 == r:
                alt, af, arv, q = a2, fwd[j, a2], rev[j, a2], q2[j]
            elif a2 == r and prob2[j] >= 0.5:
                alt, af, arv, q = a1, fwd[j, a1], rev[j, a1], q2[j]
            elif a2 != r and a1 != r and prob2[j] >= 0.5:
                f1, r1, f2, r2 = fwd[j, a1], rev[j, a1], fwd[j, a2], rev[j, a2]
                out.append(head + "%s,%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:%.4f,%.4f:%d,%d,%d:%d,%d,%d:%d,%d,%d\n" % (
                    NUM_TO_BASE[a1], NUM_TO_BASE[a2], q2[j], "PASS", info[j], "1/2", dp[j], (f1 + r1) / dp[j], (f2 + r2) / dp[j],
                    rf + rr, f1 + r1, f2 + r2, rf, f1, f2, rr, r1, r2))
                continue
            else:
                continue          # unreachable for k >= 2 (prob2 >= 0.5), kept for parity with the reference's elif chain
            out.append(head + "%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:%.4f:%d,%d:%d,%d:%d,%d\n" % (
                NUM_TO_BASE[alt], q, "PASS", info[j], "0/1", dp[j], (af + arv) / dp[j], rf + rr, af + arv, rf, af, rr, arv))
        elif k[j] == 1 and r != p1[j]:
            a1 = p1[j]
            af, arv = fwd[j, a1], rev[j, a1]
            out.append(head + "%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:%.4f:%d,%d:%d,%d:%d,%d\n" % (
                NUM_TO_BASE[a1], q1[j], "PASS", info[j], "1/1", dp[j], (af + arv) / dp[j], rf + rr, af + arv, rf, af, rr, arv))
        elif k[j] == 1:
            out.append(head + "%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:.:.:.:.\n" % (".", q1[j], "REF", info[j], "./.", dp[j]))
        else:
            out.append(head + "%s\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t%s:%d:.:.:.:.\n" % (".", 0, "LOW", info[j], "./.", dp[j]))
    return out


def call_chunks_blob(params, chunks, snp_weights, hap_weights=None, device=0, impl=0, threads=0):
    """All chunks of ONE contig on the GPU, records formatted by the library (nc_format_snp_records, threaded):
    -> (blob bytes, line_off int64[n+1], is_pass bool[n], pos int32[n]) in (chunk, position) order — what snpCaller.caller writes
    for those chunks (snpCaller.py:83-198).

    params: the reference's dict (sam_path, fasta_path, threshold, mincov, maxcov, min_allele_freq, min_nbr_sites, seq,
            supplementary, exclude_bed, disable_coverage_normalization)
    snp_weights: (tensors, train_coverage) of the diploid model; hap_weights: tensors of the haploid model."""
    empty = (b"", np.zeros(1, np.int64), np.zeros(0, bool), np.zeros(0, np.int32))
    if not chunks:
        return empty
    chrom, ploidy = chunks[0]["chrom"], chunks[0]["ploidy"]
    assert all(c["chrom"] == chrom and c["ploidy"] == ploidy for c in chunks)
    ctx = snp_pileups.context(device)
    from . import weights as W
    if ploidy == "haploid":
        if hap_weights is None:
            raise ValueError("haploid region needs the haploid SNP model")
        ctx.load_snp_weights(W.pack_snp_blob(hap_weights, True), 30.0, True)                  # snpCaller.py:73
    else:
        tensors, tc = snp_weights
        ctx.load_snp_weights(W.pack_snp_blob(tensors, False), tc, False)
    # both ploidy branches honour --disable_coverage_normalization (snpCaller.py:93-96 diploid, :169-172 haploid: hap_train_coverage / dp)
    normalize = not params.get("disable_coverage_normalization", False)
    rs = sources.resolve(params["sam_path"], chrom)
    bed = sources.bed_intervals(params.get("exclude_bed"), chrom)
    n = snp_pileups.scan_chunks(ctx, rs, params, chunks, ploidy, bed)
    if n == 0:
        return empty
    probs = ctx.snp_forward(normalize=normalize, impl=impl)
    _, meta, _, _ = ctx.snp_fetch(want_mat=False)
    blob, off, ok = capi.format_snp_records(chrom, meta["pos"], meta["ref_code"], probs, meta["dp"], meta["alt"], meta["fwd"], meta["rev"],
                                            haploid=(ploidy == "haploid"), threads=threads)
    return blob, off, ok, meta["pos"].astype(np.int32)


def call_chunks(params, chunks, snp_weights, hap_weights=None, device=0, impl=0):
    """`call_chunks_blob` as a list of VCF record lines (the unfiltered records of those chunks, in order)."""
    blob, off, _, _ = call_chunks_blob(params, chunks, snp_weights, hap_weights, device, impl)
    return [blob[off[i]:off[i + 1]].decode() for i in range(len(off) - 1) if off[i + 1] > off[i]]
