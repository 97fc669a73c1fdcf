"""VCF assembly: header lines of snpCaller.call_manager (snpCaller.py:259-276) and
indelCaller.call_manager (indelCaller.py:373-383), coordinate sort (what `bcftools sort` does to the concatenated
per-process files), PASS filter (`bcftools view -f PASS`, snpCaller.py:285) and BGZF output.  No CSI/tabix index is
written (not on the measured path)."""
from .bamio import bgzf_compress

SNP_HEADER = (
    '##fileformat=VCFv4.2\n'
    '##FILTER=<ID=PASS,Description="All filters passed">\n'
    '##FILTER=<ID=LOW,Description="All alleles have probability less than 50%.">\n'
    '##FILTER=<ID=REF,Description="Homozygous Reference. Only reference allele has greater than 50% probability. All alternative alleles having probability less than 50%.">\n'
    '{contigs}'
    '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
    '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Depth">\n'
    '##FORMAT=<ID=AD,Number=R,Type=Integer,Description="Allelic depths for the ref and alt alleles in the order listed">\n'
    '##FORMAT=<ID=ADF,Number=R,Type=Integer,Description="Allelic depths on forward strand for the ref and alt alleles in the order listed">\n'
    '##FORMAT=<ID=ADR,Number=R,Type=Integer,Description="Allelic depths on reverse strand for the ref and alt alleles in the order listed">\n'
    '##FORMAT=<ID=VF,Number=A,Type=Float,Description="Alternative allele frequency in the order listed">\n'
    '##INFO=<ID=PR,Number=4,Type=Float,Description="Probability of presence of alleles A, C, G and T, in the given order. Probability of each base is out of 1, independent of each other.">\n'
    '##INFO=<ID=FQ,Number=1,Type=Float,Description="Maximum frequency of non-reference base.">\n'
    '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t{sample}\n')

INDEL_HEADER = (
    '##fileformat=VCFv4.2\n'
    '##FILTER=<ID=PASS,Description="All filters passed">\n'
    '{contigs}'
    '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
    '##FORMAT=<ID=GQ,Number=1,Type=Float,Description="Genotype Probability">\n'
    '##FORMAT=<ID=PS,Number=1,Type=Integer,Description="Phase set identifier">\n'
    '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t{sample}\n')


def header(kind, contigs, sample="SAMPLE"):
    """`contigs` in the order to print (the reference iterates a Python set for SNPs — order unspecified).
    kind 'all' = the merged file of `--mode all` (the reference's `bcftools concat` of the SNP and indel files,
    indelCaller.py:385-395): the union of both headers."""
    ctg = "".join("##contig=<ID=%s>\n" % c for c in contigs)
    if kind == "all":
        snp = SNP_HEADER.format(contigs=ctg, sample=sample).splitlines(True)
        ind = INDEL_HEADER.format(contigs=ctg, sample=sample).splitlines(True)
        extra = [ln for ln in ind if ln.startswith("##") and ln not in snp]
        return "".join(snp[:-1] + extra + snp[-1:])
    tmpl = SNP_HEADER if kind == "snps" else INDEL_HEADER
    return tmpl.format(contigs=ctg, sample=sample)


def read_records(path):
    """Record lines (no header) of a VCF written by `write_vcf` (plain or BGZF)."""
    import gzip
    opener = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    with opener(path, "rt") as f:
        return [ln for ln in f if not ln.startswith("#")]


def sort_records(lines, contigs):
    """Stable coordinate sort by (contig order, POS): duplicates from shared chunk boundaries are both kept, like bcftools sort."""
    rank = {c: i for i, c in enumerate(contigs)}

    def key(ln):
        f = ln.split("\t", 2)
        return rank.get(f[0], len(rank)), int(f[1])
    return sorted(lines, key=key)


def pass_only(lines):
    return [ln for ln in lines if ln.split("\t", 7)[6] == "PASS"]


def write_vcf(path, kind, contigs, lines, sample="SAMPLE"):
    """Write header + sorted records; BGZF-compressed when the path ends in .gz."""
    data = (header(kind, contigs, sample) + "".join(sort_records(lines, contigs))).encode()
    with open(path, "wb") as f:
        f.write(bgzf_compress(data) if path.endswith(".gz") else data)


def write_vcf_blobs(path, kind, contigs, parts, sample="SAMPLE", pass_only=False):
    """Like `write_vcf` for records that arrive as byte blobs: parts = [(chrom, blob, line_off, is_pass, pos), ...], each in
    (chunk, position) order (what `snp_caller.call_chunks_blob` returns).  Output order = contig order, then position, stable
    (both records of a shared chunk boundary are kept in chunk order, like `sort_records`)."""
    import numpy as np
    rank = {c: i for i, c in enumerate(contigs)}
    by_contig = {}
    for part in parts:
        by_contig.setdefault(part[0], []).append(part)
    body, n_written = [], 0
    for chrom in sorted(by_contig, key=lambda c: rank.get(c, len(rank))):
        group = [p for p in by_contig[chrom] if len(p[4])]
        if not group:
            continue
        pos = np.concatenate([p[4] for p in group])
        has = np.concatenate([np.diff(p[2]) > 0 for p in group])
        keep = has & (np.concatenate([p[3] for p in group]) if pass_only else True)
        in_order = len(pos) < 2 or not np.any(np.diff(pos) < 0)
        if len(group) == 1 and in_order and keep.all():
            body.append(group[0][1])                                   # the common case: one part, already sorted, nothing dropped
            n_written += len(pos)
            continue
        part_of = np.concatenate([np.full(len(p[4]), i) for i, p in enumerate(group)])
        local = np.concatenate([np.arange(len(p[4])) for p in group])
        idx = np.arange(len(pos)) if in_order else np.argsort(pos, kind="stable")
        idx = idx[keep[idx]]
        body.append(b"".join(group[part_of[k]][1][group[part_of[k]][2][local[k]]:group[part_of[k]][2][local[k] + 1]] for k in idx.tolist()))
        n_written += len(idx)
    data = header(kind, contigs, sample).encode() + b"".join(body)
    with open(path, "wb") as f:
        f.write(bgzf_compress(data) if path.endswith(".gz") else data)
    return n_written
