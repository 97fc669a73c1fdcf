"""Indel record normalisation after the indel stage: what the reference gets from
`rtg vcfdecompose -i - -o - | rtg vcffilter --non-snps-only` (indelCaller.py:391).

RTG Tools is an external program that is not part of the reference repository and not available here, so this is this
package's own, fully stated rule set — "parity unpinned" against rtg — with the same purpose: the indel stage writes
alleles padded with the context the allele walk kept (`TACGGGGCTTCCTCA -> TACGGGGCTTC`), and downstream tools expect
minimal, anchored representations (`CCTCA -> C` ten bases further).  Rules, per record and per ALT allele:

  1. strip the common suffix, then the common prefix of REF / ALT;
  2. nothing left on one side  -> one indel, anchored on the base before it (the last stripped prefix base; if the difference
     starts at the first base of the record, the base AFTER the indel anchors it, as VCF 4.2 allows, and POS stays);
  3. something left on both sides -> global affine alignment of the two remainders (nc_nw_trace, gap open 9, extend 1,
     match 20, mismatch -10 — the scores the allele walk uses, generate_indel_pileups.py:79) and one component per maximal run
     of one kind: a gap gives an anchored indel (adjacent insertion and deletion ops stay one component), a run of >= 2
     mismatches an MNP, a single mismatch a SNP (a remainder whose alignment opens with a gap at the record's first base has no
     anchor and stays one complex record);
  4. `--non-snps-only`: SNP components are dropped, indels and MNPs stay;
  5. genotypes: a component of allele k of a `1|2` (`1/2`) call becomes `1|0` (first allele) or `0|1` (second); the same
     component from both alleles merges into `1|1`, two different ALTs on the same POS and REF merge back into `1|2`;
     every other GT is copied.  QUAL, FILTER, INFO and the remaining FORMAT fields are copied.
Records come back coordinate-sorted per contig in input order of contigs."""
import numpy as np

from . import capi

_CODE = {"A": 0, "G": 1, "T": 2, "C": 3}


def _codes(s):
    return np.array([_CODE.get(c, 4) for c in s], np.uint8)


def components(pos, ref, alt):
    """-> list of (pos, ref, alt, kind) with kind in 'indel' | 'mnp' | 'snp' | 'complex' for one REF / ALT pair (rules 1-3)."""
    if ref == alt:
        return []
    suf = 0
    while suf < min(len(ref), len(alt)) and ref[len(ref) - 1 - suf] == alt[len(alt) - 1 - suf]:
        suf += 1
    r, a = ref[:len(ref) - suf], alt[:len(alt) - suf]
    pre = 0
    while pre < min(len(r), len(a)) and r[pre] == a[pre]:
        pre += 1
    r2, a2 = r[pre:], a[pre:]
    if not r2 or not a2:                                     # pure insertion / deletion
        if pre > 0:
            anchor = ref[pre - 1]
            return [(pos + pre - 1, anchor + r2, anchor + a2, "indel")]
        nxt = ref[len(r)] if len(r) < len(ref) else ""        # difference at the very first base: anchor on the base after it
        return [(pos, r2 + nxt, a2 + nxt, "indel")] if nxt else [(pos, ref, alt, "indel")]
    if len(r2) == len(a2) and all(x != y for x, y in zip(r2, a2)):
        return [(pos + pre, r2, a2, "snp" if len(r2) == 1 else "mnp")]
    out = []
    i = j = 0                                                # i over a2 (query), j over r2 (reference)
    cig = capi.nw_trace(_codes(a2), _codes(r2), 9, 1, 20, -10)
    if pre == 0 and cig and cig[0][0] in (1, 2):             # a gap with no base before it inside a complex difference: keep it whole
        return [(pos, r2, a2, "complex")]
    k = 0
    while k < len(cig):
        op, ln = cig[k]
        if op == 7:
            i += ln; j += ln; k += 1
        elif op == 8:
            out.append((pos + pre + j, r2[j:j + ln], a2[i:i + ln], "snp" if ln == 1 else "mnp"))
            i += ln; j += ln; k += 1
        else:                                                # adjacent insertion / deletion ops form one component
            ins = dele = ""
            while k < len(cig) and cig[k][0] in (1, 2):
                if cig[k][0] == 1:
                    ins += a2[i:i + cig[k][1]]; i += cig[k][1]
                else:
                    dele += r2[j:j + cig[k][1]]; j += cig[k][1]
                k += 1
            at = pre + j - len(dele)                         # offset of the first affected reference base; > 0 here
            anchor = ref[at - 1]
            out.append((pos + at - 1, anchor + dele, anchor + ins, "indel"))
    return out


def decompose_records(lines, contigs=None, keep_snps=False):
    """Record lines of the indel stage -> normalised record lines (rules 1-5)."""
    rank = {}
    for ln in lines:
        rank.setdefault(ln.split("\t", 1)[0], len(rank))
    if contigs:
        rank = {c: i for i, c in enumerate(contigs)}
    out = []                                                 # (contig rank, pos, input order, emit order, fields, ref, alt, gt)
    for n, ln in enumerate(lines):
        f = ln.rstrip("\n").split("\t")
        pos, ref, alts = int(f[1]), f[3], f[4].split(",")
        gt = f[9].split(":", 1)[0]
        sep = "|" if "|" in gt else "/"
        idx = gt.replace("|", "/").split("/")
        two = len(alts) == 2 and sorted(idx) == ["1", "2"]
        found = {}                                           # (pos, ref, alt) -> haplotypes (0 left of the bar, 1 right) carrying it
        for k, alt in enumerate(alts):
            hap = idx.index(str(k + 1)) if two else None
            for p, r, a, kind in components(pos, ref, alt):
                if kind == "snp" and not keep_snps:
                    continue
                found.setdefault((p, r, a), set()).add(hap)
        emit = []                                            # [pos, ref, alt, gt] in order of first appearance
        if not two:
            emit = [[p, r, a, gt] for (p, r, a) in found]
        else:
            by_site = {}
            for (p, r, a), haps in found.items():
                if len(haps) == 2:
                    emit.append([p, r, a, "1" + sep + "1"])
                    continue
                h = next(iter(haps))
                mate = by_site.get((p, r))
                if mate is not None and mate[4] != h:        # two different ALTs on the same POS and REF, one per haplotype: 1|2 again
                    first, second = (mate[2], a) if mate[4] == 0 else (a, mate[2])
                    mate[2], mate[3] = first + "," + second, "1" + sep + "2"
                    continue
                rec = [p, r, a, ("1" + sep + "0") if h == 0 else ("0" + sep + "1"), h]
                by_site.setdefault((p, r), rec)
                emit.append(rec)
        for e, rec in enumerate(emit):
            out.append((rank.get(f[0], len(rank)), rec[0], n, e, f, rec[1], rec[2], rec[3]))
    out.sort(key=lambda t: t[:4])
    res = []
    for _, p, _, _, f, r, a, g in out:
        h = list(f)
        h[1], h[3], h[4] = str(p), r, a
        h[9] = ":".join([g] + f[9].split(":")[1:])
        res.append("\t".join(h) + "\n")
    return res
