"""VCF assembly: header lines of snpCaller.call_manager (snpCaller.py:259-276) and
indelCaller.call_manager (indelCaller.py:373-383), coordinate sort (what `bcftools sort` does to the concatenated
per-process files), PASS filter (`bcftools view -f PASS`, snpCaller.py:285), BGZF output and the CSI index next to it
(`tabix -fp vcf --csi`, snpCaller.py:283-285)."""
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
    if kind == "phased_snps":                    # the SNP header plus the PS line the phasing step adds (what `whatshap phase` writes, indelCaller.py:237)
        snp = SNP_HEADER.format(contigs=ctg, sample=sample).splitlines(True)
        return "".join(snp[:-1] + ['##FORMAT=<ID=PS,Number=1,Type=Integer,Description="Phase set identifier">\n'] + snp[-1:])
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


def _write(path, data, index):
    if path.endswith(".gz") and index:
        write_indexed(path, data)
    else:
        with open(path, "wb") as f:
            f.write(bgzf_compress(data) if path.endswith(".gz") else data)


def write_vcf(path, kind, contigs, lines, sample="SAMPLE", index=False):
    """Write header + sorted records; BGZF-compressed when the path ends in .gz, with a CSI index next to it when `index`."""
    data = (header(kind, contigs, sample) + "".join(sort_records(lines, contigs))).encode()
    _write(path, data, index)


def write_vcf_blobs(path, kind, contigs, parts, sample="SAMPLE", pass_only=False, index=False):
    """Like `write_vcf` for records that arrive as byte blobs: parts = [(chrom, blob, line_off, is_pass, pos), ...], each in
    (chunk, position) order (what `snp_caller.call_chunks_blob` returns; single-base REF alleles).  Output order = contig order,
    then position, stable (both records of a shared chunk boundary are kept in chunk order, like `sort_records`)."""
    import numpy as np
    rank = {c: i for i, c in enumerate(contigs)}
    names = list(contigs)
    by_contig = {}
    for part in parts:
        by_contig.setdefault(part[0], []).append(part)
    body, n_written = [], 0
    rec_rid, rec_pos, rec_len = [], [], []
    for chrom in sorted(by_contig, key=lambda c: rank.get(c, len(rank))):
        group = [p for p in by_contig[chrom] if len(p[4])]
        if not group:
            continue
        if chrom not in rank:
            rank[chrom] = len(names)
            names.append(chrom)
        pos = np.concatenate([p[4] for p in group])
        lens = np.concatenate([np.diff(p[2]) for p in group])
        keep = (lens > 0) & (np.concatenate([p[3] for p in group]) if pass_only else True)
        in_order = len(pos) < 2 or not np.any(np.diff(pos) < 0)
        if len(group) == 1 and in_order and keep.all():
            body.append(group[0][1])                                   # the common case: one part, already sorted, nothing dropped
            idx = np.arange(len(pos))
        else:
            part_of = np.concatenate([np.full(len(p[4]), i) for i, p in enumerate(group)])
            local = np.concatenate([np.arange(len(p[4])) for p in group])
            idx = np.arange(len(pos)) if in_order else np.argsort(pos, kind="stable")
            idx = idx[keep[idx]]
            body.append(b"".join(group[part_of[k]][1][group[part_of[k]][2][local[k]]:group[part_of[k]][2][local[k] + 1]] for k in idx.tolist()))
        n_written += len(idx)
        rec_rid.append(np.full(len(idx), rank[chrom], np.int64)); rec_pos.append(pos[idx].astype(np.int64)); rec_len.append(lens[idx].astype(np.int64))
    head = header(kind, contigs, sample).encode()
    data = head + b"".join(body)
    if path.endswith(".gz") and index:
        rid = np.concatenate(rec_rid) if rec_rid else np.zeros(0, np.int64)
        p1 = np.concatenate(rec_pos) if rec_pos else np.zeros(0, np.int64)
        ln = np.concatenate(rec_len) if rec_len else np.zeros(0, np.int64)
        u0 = len(head) + np.concatenate([[0], np.cumsum(ln)[:-1]]) if len(ln) else np.zeros(0, np.int64)
        write_indexed(path, data, records=(names, rid, p1 - 1, p1, u0, ln))
    else:
        _write(path, data, False)
    return n_written


# ------------------------------------------------------------------------------------------------ CSI index (tabix -p vcf --csi)
def _reg2bin(beg, end, min_shift=14, depth=5):
    """Bin of the 0-based half-open interval [beg, end) in the CSI binning scheme (CSIv1 specification)."""
    end -= 1
    s, t = min_shift, ((1 << depth * 3) - 1) // 7
    for l in range(depth, 0, -1):
        if beg >> s == end >> s:
            return t + (beg >> s)
        s += 3
        t -= 1 << (l - 1) * 3       # t -= 1 << ((l-1)*3): offset of the next coarser level
    return 0


def _bin_first_window(b, depth=5):
    """Index of the first min_shift-sized window covered by bin b (htslib hts_bin_bot)."""
    l, bb = 0, b
    while bb:
        l += 1
        bb = (bb - 1) >> 3
    first_of_level = ((1 << 3 * l) - 1) // 7
    return (b - first_of_level) << (depth - l) * 3


def _reg2bin_np(beg, end, min_shift=14, depth=5):
    import numpy as np
    beg = np.asarray(beg, np.int64)
    e1 = np.asarray(end, np.int64) - 1
    out = np.zeros(len(beg), np.int64)
    done = np.zeros(len(beg), bool)
    s, t = min_shift, ((1 << depth * 3) - 1) // 7
    for l in range(depth, 0, -1):
        hit = ~done & ((beg >> s) == (e1 >> s))
        out[hit] = t + (beg[hit] >> s)
        done |= hit
        s += 3
        t -= 1 << (l - 1) * 3
    return out


def build_csi_arrays(names, rid, beg, end, u0, nbytes, block_sizes, block_bytes=0xff00, min_shift=14, depth=5):
    """CSI index bytes from per-record arrays in FILE ORDER: reference id, 0-based [beg, end), byte offset of the line in the
    uncompressed text and its length.  Vectorised: chunks are runs of consecutive records of one bin."""
    import struct
    import numpy as np
    rid = np.asarray(rid, np.int64); beg = np.asarray(beg, np.int64); end = np.asarray(end, np.int64)
    u0 = np.asarray(u0, np.int64); u1 = u0 + np.asarray(nbytes, np.int64)
    coff = np.concatenate([[0], np.cumsum(block_sizes)]).astype(np.int64)
    v0 = (coff[u0 // block_bytes] << 16) | (u0 % block_bytes)
    v1 = (coff[u1 // block_bytes] << 16) | (u1 % block_bytes)
    bins = _reg2bin_np(beg, end, min_shift, depth)
    nm = b"".join(c.encode() + b"\0" for c in names)
    aux = struct.pack("<iiiiiii", 2, 1, 2, 0, ord("#"), 0, len(nm)) + nm          # format VCF, col_seq 1, col_beg 2, col_end 0, meta '#', skip 0
    out = [b"CSI\x01", struct.pack("<iii", min_shift, depth, len(aux)), aux, struct.pack("<i", len(names))]
    for i in range(len(names)):
        sel = np.flatnonzero(rid == i)
        if len(sel) == 0:
            out.append(struct.pack("<i", 0))
            continue
        b_i, s0, s1 = bins[sel], v0[sel], v1[sel]
        cut = np.flatnonzero(np.diff(b_i) != 0) + 1
        starts = np.concatenate([[0], cut]); stops = np.concatenate([cut, [len(sel)]])
        table = {}
        for a_, z_ in zip(starts.tolist(), stops.tolist()):
            ch = table.setdefault(int(b_i[a_]), [])
            lo, hi = int(s0[a_]), int(s1[z_ - 1])
            if ch and ch[-1][1] == lo:
                ch[-1][1] = hi
            else:
                ch.append([lo, hi])
        # linear index: smallest virtual offset of a record overlapping each 2^min_shift window; empty windows inherit the next one
        wb, we = beg[sel] >> min_shift, (end[sel] - 1) >> min_shift
        nwin = int(we.max()) + 1
        lin = np.full(nwin, -1, np.int64)
        first_w, first_i = np.unique(wb, return_index=True)
        lin[first_w] = s0[first_i]
        for k in np.flatnonzero(we > wb).tolist():                  # records reaching into further windows (long REF alleles)
            for w in range(int(wb[k]) + 1, int(we[k]) + 1):
                if lin[w] < 0 or s0[k] < lin[w]:
                    lin[w] = s0[k]
        nxt = 0
        loff = lin.tolist()
        for w in range(nwin - 1, -1, -1):
            if loff[w] >= 0:
                nxt = loff[w]
            else:
                loff[w] = nxt
        out.append(struct.pack("<i", len(table)))
        for b in sorted(table):
            w0 = _bin_first_window(b, depth)
            out.append(struct.pack("<IQi", b, loff[w0] if w0 < nwin else 0, len(table[b])))
            for lo, hi in table[b]:
                out.append(struct.pack("<QQ", lo, hi))
    out.append(struct.pack("<Q", 0))
    return b"".join(out)


def build_csi(data, contigs, block_bytes=0xff00, block_sizes=None, min_shift=14, depth=5):
    """CSI index (coordinate-sorted index, CSIv1 + the tabix header in its auxiliary block) of the VCF text `data` as it is written
    in BGZF blocks of `block_bytes` uncompressed bytes — what the reference produces with `tabix -fp vcf --csi`
    (snpCaller.py:283-285, indelCaller.py:396).  `block_sizes`: compressed size of every block.  Returns the index bytes."""
    names = list(contigs)
    ids = {c: i for i, c in enumerate(names)}
    rid, beg, end, u0, nb = [], [], [], [], []
    u = 0
    for ln in data.split(b"\n"):
        n = len(ln) + 1
        if ln and not ln.startswith(b"#"):
            f = ln.split(b"\t", 4)
            chrom = f[0].decode()
            if chrom not in ids:
                ids[chrom] = len(names)
                names.append(chrom)
            p0 = int(f[1]) - 1
            rid.append(ids[chrom]); beg.append(p0); end.append(p0 + max(1, len(f[3]))); u0.append(u); nb.append(n)
        u += n
    return build_csi_arrays(names, rid, beg, end, u0, nb, block_sizes, block_bytes, min_shift, depth)


def write_indexed(path, data, records=None):
    """Write `data` as BGZF to `path` and its CSI index to `path + ".csi"` (both BGZF, like bgzip + tabix --csi).
    records = (names, rid, beg, end, u0, nbytes) spares the text parse when the caller knows the records (SNP blobs)."""
    from .bamio import _BGZF_EOF, _bgzf_block
    contigs = [ln[len("##contig=<ID="):-1].split(",")[0] for ln in data.split(b"#CHROM", 1)[0].decode().splitlines() if ln.startswith("##contig=<ID=")]
    chunks = [data[off:off + 0xff00] for off in range(0, len(data), 0xff00)]
    if len(chunks) >= 64:
        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(min(32, os.cpu_count() or 1)) as ex:
            blocks = list(ex.map(lambda c: _bgzf_block(c, 4), chunks, chunksize=16))
    else:
        blocks = [_bgzf_block(c, 4) for c in chunks]
    with open(path, "wb") as f:
        f.write(b"".join(blocks) + _BGZF_EOF)
    sizes = [len(b) for b in blocks]
    idx = build_csi_arrays(*records, block_sizes=sizes) if records is not None else build_csi(data, contigs, block_sizes=sizes)
    with open(path + ".csi", "wb") as f:
        f.write(bgzf_compress(idx))


class _BgzfText:
    """Random access into a BGZF text file by virtual offset (compressed block start << 16 | offset inside the block)."""

    def __init__(self, path):
        self.raw = open(path, "rb").read()
        self.cache = {}

    def block(self, co):
        if co not in self.cache:
            import struct
            import zlib
            xlen = struct.unpack_from("<H", self.raw, co + 10)[0]
            bsize = struct.unpack_from("<H", self.raw, co + 16)[0] + 1
            self.cache[co] = (zlib.decompress(self.raw[co + 12 + xlen:co + bsize - 8], -15), co + bsize)
        return self.cache[co]

    def read_line(self, v):
        """-> (line including its newline, virtual offset of the next line)."""
        co, uo = v >> 16, v & 0xFFFF
        parts = []
        while True:
            data, nxt = self.block(co)
            e = data.find(b"\n", uo)
            if e >= 0:
                parts.append(data[uo:e + 1])
                uo = e + 1
                if uo == len(data):
                    co, uo = nxt, 0
                return b"".join(parts), (co << 16) | uo
            parts.append(data[uo:])
            co, uo = nxt, 0


def csi_query(path, chrom, beg, end):
    """Records of a BGZF VCF overlapping the 0-based half-open region [beg, end), fetched through its CSI index the way tabix does:
    bins overlapping the region -> their chunks -> lines between the chunks' virtual offsets."""
    import gzip
    import struct
    idx = gzip.open(path + ".csi", "rb").read()
    if idx[:4] != b"CSI\x01":
        raise ValueError("not a CSI index")
    min_shift, depth, l_aux = struct.unpack_from("<iii", idx, 4)
    aux = idx[16:16 + l_aux]
    l_nm = struct.unpack_from("<i", aux, 24)[0]
    names = aux[28:28 + l_nm].split(b"\0")[:-1]
    p = 16 + l_aux
    n_ref = struct.unpack_from("<i", idx, p)[0]
    p += 4
    want = names.index(chrom.encode()) if chrom.encode() in names else -1
    table = {}
    for i in range(n_ref):
        n_bin = struct.unpack_from("<i", idx, p)[0]
        p += 4
        for _ in range(n_bin):
            b, loff, n_chunk = struct.unpack_from("<IQi", idx, p)
            p += 16
            if i == want:
                table[b] = (loff, [struct.unpack_from("<QQ", idx, p + 16 * k) for k in range(n_chunk)])
            p += 16 * n_chunk
    if want < 0 or end <= beg:
        return []
    cand, s, t = [0], min_shift + depth * 3, 0
    for l in range(1, depth + 1):
        s -= 3
        t += 1 << (l - 1) * 3
        cand += list(range(t + (beg >> s), t + ((end - 1) >> s) + 1))
    # the finest overlapping bin's loffset bounds the chunks from below (records before it cannot overlap the region)
    w0 = beg >> min_shift
    fine = ((1 << depth * 3) - 1) // 7 + w0
    min_off = table[fine][0] if fine in table else 0
    chunks = sorted(c for b in cand if b in table for c in table[b][1] if c[1] > min_off)
    rd = _BgzfText(path)
    out, seen = [], set()
    for v0, v1 in chunks:
        v = v0
        while v < v1:
            at = v
            ln, v = rd.read_line(v)
            f = ln.split(b"\t", 4)
            rb = int(f[1]) - 1
            if at not in seen and f[0] == chrom.encode() and rb < end and rb + max(1, len(f[3])) > beg:
                seen.add(at)
                out.append((at, ln.decode()))
    return [ln for _, ln in sorted(out)]
