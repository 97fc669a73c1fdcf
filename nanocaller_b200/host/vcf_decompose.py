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
    out = []                                                 # (contig rank, pos, seq no, fields)
    rank = {}
    for ln in lines:
        f = ln.rstrip("\n").split("\t")
        rank.setdefault(f[0], len(rank))
    if contigs:
        rank = {c: i for i, c in enumerate(contigs)}
    for n, ln in enumerate(lines):
        f = ln.rstrip("\n").split("\t")
        pos, ref, alts = int(f[1]), f[3], f[4].split(",")
        sample = f[9].split(":")
        gt = sample[0]
        sep = "|" if "|" in gt else "/"
        two = len(alts) == 2 and sorted(gt.replace("|", "/").split("/")) == ["1", "2"]
        comps = {}                                           # (pos, ref, alt) -> set of allele indices carrying it
        order = []
        for k, alt in enumerate(alts):
            for (p, r, a, kind) in components(pos, ref, alt):
                if kind == "snp" and not keep_snps:
                    continue
                key = (p, r, a)
                if key not in comps:
                    comps[key] = set(); order.append(key)
                comps[key].add(k)
        first_is_1 = gt.replace("|", "/").split("/")[0] == "1"
        merged = {}                                          # same POS + REF with different ALTs from the two alleles -> one 1|2 record
        for key in order:
            p, r, a = key
            if two:
                ks = comps[key]
                if len(ks) == 2:
                    g = "1" + sep + "1"
                else:
                    k = next(iter(ks))
                    on_first = (k == 0) == first_is_1        # allele index 0 is ALT '1'
                    g = ("1" + sep + "0") if on_first else ("0" + sep + "1")
                    other = merged.get((p, r))
                    if other is not None and other[2] != g and len(other[3]) == 1 and "1" + sep + "1" not in (other[2], g):
                        a1, a2_ = (other[1], a) if other[2].startswith("1") else (a, other[1])
                        other[1], other[2] = a1 + "," + a2_, "1" + sep + "2"
                        other[3].append(k)
                        continue
                rec = [p, a, g, [0]]
                merged.setdefault((p, r), rec)
                out.append((rank.get(f[0], len(rank)), p, n, f, r, rec))
            else:
                rec = [p, a, gt, [0]]
                out.append((rank.get(f[0], len(rank)), p, n, f, r, rec))
    out.sort(key=lambda t: (t[0], t[1], t[2]))
    res = []
    for _, p, _, f, r, rec in out:
        g = list(f)
        g[1], g[3], g[4] = str(p), r, rec[1]
        g[9] = ":".join([rec[2]] + f[9].split(":")[1:])
        res.append("\t".join(g) + "\n")
    return res
