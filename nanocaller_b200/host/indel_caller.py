"""Host mirror of the per-chunk body of indelCaller.indel_run (indelCaller.py:66-182): candidate tensors ->
hstack of the three groups -> indel CNN -> genotype decision -> VCF record lines.

`records_from_calls` / `haploid_records_from_calls` are the decision + formatting step (G2); `call_chunk` is the
product path for one chunk (scan, build, CNN on the GPU; records on the host).  The `prev` overlap suppression makes
the decision sequential in position, exactly as in the reference (:93,:104)."""
import math

import numpy as np

from . import indel_pileups, snp_pileups, sources


def _q10(x):
    return -10.0 * math.log10(x)


def records_from_calls(chrom, pos, probs, alleles_seq, phase):
    """indelCaller.py:88-152.  probs float32 [N,4] (hom-ref, hom-alt, het-ref, het-alt, :14)."""
    probs = np.asarray(probs, np.float32)
    pred = np.argmax(probs, axis=1)
    p64 = probs.astype(np.float64)
    out = []
    prev = 0
    for j in range(len(pred)):
        if not pos[j] > prev or not probs[j, 0] <= 0.95:
            continue
        q = _q10(1e-6 + p64[j, 0])
        a0, a1, at = alleles_seq[j]
        if pred[j] == 1 and at[0]:
            out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ\t1/1:%.2f\n' % (chrom, pos[j], at[0], at[1], q, _q10(1 + 1e-6 - p64[j, 1])))
            prev = pos[j] + max(len(at[0]), len(at[1]))
        elif a0[0] and a1[0]:
            if a0[0] == a1[0] and a0[1] == a1[1]:
                out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ\t1/1:%.2f\n' % (chrom, pos[j], a0[0], a0[1], q, _q10(1 + 1e-6 - p64[j, 1])))
                prev = pos[j] + max(len(a0[0]), len(a0[1]))
            else:
                (ref1, alt1), (ref2, alt2) = a0, a1
                l = min(len(ref1), len(ref2))
                if len(ref1) > len(ref2):
                    ref, alt2 = ref1, alt2 + ref1[l:]
                else:
                    ref, alt1 = ref2, alt1 + ref2[l:]
                gq = _q10(1 + 1e-6 - p64[j, 3])
                if phase[j]:
                    out.append('%s\t%d\t.\t%s\t%s,%s\t%.2f\tPASS\t.\tGT:GQ:PS\t1|2:%.2f:%d\n' % (chrom, pos[j], ref, alt1, alt2, q, gq, phase[j]))
                else:
                    out.append('%s\t%d\t.\t%s\t%s,%s\t%.2f\tPASS\t.\tGT:GQ\t1|2:%.2f\n' % (chrom, pos[j], ref, alt1, alt2, q, gq))
                prev = pos[j] + max(len(ref), len(alt1), len(alt2))
        elif a0[0] or a1[0]:
            a, gt = (a0, "0|1") if a0[0] else (a1, "1|0")
            gq = _q10(1 + 1e-6 - p64[j, 2])
            if phase[j]:
                out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ:PS\t%s:%.2f:%d\n' % (chrom, pos[j], a[0], a[1], q, gt, gq, phase[j]))
            else:
                out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ\t%s:%.2f\n' % (chrom, pos[j], a[0], a[1], q, gt, gq))
            prev = pos[j] + max(len(a[0]), len(a[1]))
    return out


def haploid_records_from_calls(chrom, pos, probs, alleles_seq):
    """indelCaller.py:173-179."""
    p = np.asarray(probs, np.float32).reshape(-1)
    out = []
    prev = 0
    for j in range(len(p)):
        a = alleles_seq[j]
        if pos[j] > prev and p[j] >= 0.5 and a[0]:
            q = -100.0 * math.log10(1e-6 + 1 - float(p[j]))
            out.append('%s\t%d\t.\t%s\t%s\t%.2f\tPASS\t.\tGT:GQ\t1/1:%.2f\n' % (chrom, pos[j], a[0], a[1], q, q))
            prev = pos[j] + max(len(a[0]), len(a[1]))
    return out


def call_chunk(params, chunk, indel_tensors, hap_tensors=None, device=0, impl=1):
    """One chunk, like one job of indelCaller.indel_run: -> list of VCF record lines."""
    from . import weights as W
    ctx = snp_pileups.context(device)
    rs = sources.resolve(chunk["sam_path"], chunk["chrom"])
    bed = sources.bed_intervals(params.get("exclude_bed"), chunk["chrom"])
    if chunk["ploidy"] == "haploid":
        pos, x, alleles = indel_pileups.candidates_for_chunks(ctx, rs, params, [chunk], bed, haploid=True)[0]
        if len(pos) == 0:
            return []
        ctx.load_indel_weights(W.pack_indel_blob(hap_tensors), True)
        probs = ctx.indel_model_forward(np.asarray(x, np.float32), haploid=True, impl=impl)
        return haploid_records_from_calls(chunk["chrom"], pos, probs, alleles)
    pos, x0, x1, x2, alleles, phase = indel_pileups.candidates_for_chunks(ctx, rs, params, [chunk], bed)[0]
    if len(pos) == 0:
        return []
    ctx.load_indel_weights(W.pack_indel_blob(indel_tensors), False)
    x = np.hstack([x0, x1, x2]).astype(np.float32)                       # indelCaller.py:83
    probs = ctx.indel_model_forward(x, haploid=False, impl=impl)
    return records_from_calls(chunk["chrom"], pos, probs, alleles, phase)


def call_chunks(params, chunks, indel_tensors, hap_tensors=None, device=0, impl=0):
    """All chunks of ONE contig and ploidy: one scan + build on the GPU, the indel CNN on the tensors where they lie (they never
    leave the device), alleles by one batched library call; the record decision stays per chunk (the reference's `prev` overlap
    suppression restarts with every chunk, indelCaller.py:88-93).  Same lines as `call_chunk` chunk by chunk."""
    from . import weights as W
    if not chunks:
        return []
    chrom, ploidy = chunks[0]["chrom"], chunks[0]["ploidy"]
    assert all(c["chrom"] == chrom and c["ploidy"] == ploidy for c in chunks)
    ctx = snp_pileups.context(device)
    rs = sources.resolve(chunks[0]["sam_path"], chrom)
    bed = sources.bed_intervals(params.get("exclude_bed"), chrom)
    hap = ploidy == "haploid"
    meta, _, cns = indel_pileups.scan_build(ctx, rs, params, chunks, bed, haploid=hap, want_tensors=False)
    if len(meta) == 0 or not indel_pileups.kept_sites(meta, hap).any():
        return []
    ctx.load_indel_weights(W.pack_indel_blob(hap_tensors if hap else indel_tensors), hap)
    probs = ctx.indel_forward(impl=impl)                                   # a row per built site; only the kept ones are read
    pred = indel_pileups.AllelePredictions(rs, params, meta, cns, hap, device_lengths=ctx.indel_fetch_alleles()).strings()
    out = []
    for sel, pos, alleles, phase in indel_pileups.per_chunk_calls(meta, pred, len(chunks), hap):
        if len(sel) == 0:
            continue
        out += haploid_records_from_calls(chrom, pos, probs[sel], alleles) if hap else records_from_calls(chrom, pos, probs[sel], alleles, phase)
    return out
