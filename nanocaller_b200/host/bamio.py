"""BAM / BGZF / FASTA readers and writers producing the BAM-native arrays the device stager consumes
(replaces pysam.Samfile / FastaFile opening at generate_SNP_pileups.py:134-137).

Pure host code (zlib only): BGZF is a series of gzip members, BAM records are parsed into one `ReadSet` per contig with
CIGAR words and 4-bit bases copied verbatim (they already are the staging format), HP / PS integer tags extracted for
the indel path.  The writer exists so tests and the synthetic generator can produce real files."""
import gzip
import os
import struct
import zlib

import numpy as np

from .readset import ReadSet

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(chunk, level):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = co.compress(chunk) + co.flush()
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(body) + 25) + body +
            struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))


def bgzf_compress(data, level=4, threads=None):
    """bytes -> BGZF stream (64 KiB blocks with the BC extra field, plus the EOF marker).  Blocks are independent gzip members:
    large inputs are deflated on a thread pool (zlib releases the GIL)."""
    chunks = [data[off:off + 0xff00] for off in range(0, len(data), 0xff00)]
    if len(chunks) >= 64:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads or min(32, os.cpu_count() or 1)) as ex:
            out = list(ex.map(lambda c: _bgzf_block(c, level), chunks, chunksize=16))
    else:
        out = [_bgzf_block(c, level) for c in chunks]
    out.append(_BGZF_EOF)
    return b"".join(out)


def bgzf_decompress(path):
    with gzip.open(path, "rb") as f:          # gzip handles concatenated members
        return f.read()


def write_bam(path, readsets, header_text=None, index=False):
    """Write coordinate-sorted ReadSets (one per contig) as a BAM file.  qual is written as 0xff (absent).
    index=True also writes `path + ".bai"` (bins, chunks and the 16 kb linear index of the SAM specification)."""
    text = header_text or ("@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (rs.chrom, rs.contig_len) for rs in readsets))
    parts = [b"BAM\x01", struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(readsets))]
    for rs in readsets:
        nm = rs.chrom.encode() + b"\0"
        parts += [struct.pack("<i", len(nm)), nm, struct.pack("<i", rs.contig_len)]
    u = sum(len(p) for p in parts)
    recs = []                                                # (rid, beg, end, bin, u_start, u_end)
    for rid, rs in enumerate(readsets):
        ends = rs.ref_end
        for i in range(rs.n):
            name = rs.qname(i).encode() + b"\0"
            cig = rs.cigar[rs.cigar_off[i]:rs.cigar_off[i + 1]]
            seq = rs.seq4[rs.seq_off[i]:rs.seq_off[i + 1]]
            l_seq = int(rs.l_seq[i])
            aux = b""
            if rs.hp[i] > 0:
                aux = b"HPC" + struct.pack("<B", int(rs.hp[i])) + b"PSi" + struct.pack("<i", int(rs.ps[i]))
            pos = int(rs.pos[i])
            end = max(pos + 1, int(ends[i]))
            bin_ = _reg2bin(pos, end)
            cig_field = cig.astype("<u4")
            if len(cig) > 65535:         # SAM spec 4.2.2: placeholder `<l_seq>S<ref_len>N` + the real operations in CG:B,I
                aux += b"CGBI" + struct.pack("<i", len(cig)) + cig_field.tobytes()
                cig_field = np.array([(l_seq << 4) | 4, ((end - pos) << 4) | 3], "<u4")
            core = struct.pack("<iiBBHHHiiii", rid, pos, len(name), 60, bin_, len(cig_field), int(rs.flag[i]), l_seq, -1, -1, 0)
            body = core + name + cig_field.tobytes() + seq.tobytes()[:(l_seq + 1) // 2] + b"\xff" * l_seq + aux
            rec = struct.pack("<i", len(body)) + body
            parts.append(rec)
            recs.append((rid, pos, end, bin_, u, u + len(rec)))
            u += len(rec)
    stream = b"".join(parts)
    blocks = [_bgzf_block(stream[off:off + 0xff00], 4) for off in range(0, len(stream), 0xff00)]
    with open(path, "wb") as f:
        f.write(b"".join(blocks) + _BGZF_EOF)
    if not index:
        return
    coff = np.concatenate([[0], np.cumsum([len(b) for b in blocks])]).astype(np.int64)

    def voff(x):
        return (int(coff[x // 0xff00]) << 16) | (x % 0xff00)
    out = [b"BAI\x01", struct.pack("<i", len(readsets))]
    for rid, rs in enumerate(readsets):
        bins, linear = {}, {}
        for r, beg, end, bin_, u0, u1 in recs:
            if r != rid:
                continue
            v0, v1 = voff(u0), voff(u1)
            ch = bins.setdefault(bin_, [])
            if ch and ch[-1][1] == v0:
                ch[-1][1] = v1
            else:
                ch.append([v0, v1])
            for w in range(beg >> 14, ((end - 1) >> 14) + 1):
                linear[w] = min(linear.get(w, v0), v0)
        out.append(struct.pack("<i", len(bins)))
        for bin_ in sorted(bins):
            out.append(struct.pack("<Ii", bin_, len(bins[bin_])))
            for v0, v1 in bins[bin_]:
                out.append(struct.pack("<QQ", v0, v1))
        n_intv = (max(linear) + 1) if linear else 0
        out.append(struct.pack("<i", n_intv))
        last = 0
        for w in range(n_intv):
            last = linear.get(w, last)
            out.append(struct.pack("<Q", last))
    with open(path + ".bai", "wb") as f:
        f.write(b"".join(out))


def _reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


_AUX_FIXED = {ord("A"): 1, ord("c"): 1, ord("C"): 1, ord("s"): 2, ord("S"): 2, ord("i"): 4, ord("I"): 4, ord("f"): 4}
_AUX_FMT = {ord("c"): "<b", ord("C"): "<B", ord("s"): "<h", ord("S"): "<H", ord("i"): "<i", ord("I"): "<I"}


def _aux_ints(buf, off, end, want=(b"HP", b"PS")):
    """Integer values of the wanted aux tags of one record; a CG:B,I array (the real CIGAR of a read with more than 65,535
    operations, SAM spec 4.2.2) is returned under b"CG" as a uint32 array."""
    out = {}
    while off + 3 <= end:
        tag, typ = bytes(buf[off:off + 2]), buf[off + 2]
        off += 3
        if typ in _AUX_FIXED:
            if tag in want and typ in _AUX_FMT:
                out[tag] = struct.unpack_from(_AUX_FMT[typ], buf, off)[0]
            off += _AUX_FIXED[typ]
        elif typ in (ord("Z"), ord("H")):
            off = buf.index(b"\0", off) + 1
        elif typ == ord("B"):
            sub = buf[off]
            n = struct.unpack_from("<i", buf, off + 1)[0]
            if tag == b"CG" and sub == ord("I"):
                out[b"CG"] = np.frombuffer(buf, "<u4", n, off + 5)
            off += 5 + n * _AUX_FIXED.get(sub, 1)
        else:
            break
    return out


def read_bam(path, fasta=None, contigs=None):
    """-> (list of ReadSet in header order, header text).  `fasta`: {contig: uint8 array}; contigs without a reference
    get an all-'N' reference.  Unmapped records (refID < 0) are dropped.  Records must be coordinate-sorted."""
    buf = bgzf_decompress(path)
    if buf[:4] != b"BAM\x01":
        raise ValueError("%s: not a BAM file" % path)
    l_text = struct.unpack_from("<i", buf, 4)[0]
    text = buf[8:8 + l_text].decode(errors="replace")
    off = 8 + l_text
    n_ref = struct.unpack_from("<i", buf, off)[0]
    off += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", buf, off)[0]
        name = buf[off + 4:off + 4 + ln - 1].decode()
        lref = struct.unpack_from("<i", buf, off + 4 + ln)[0]
        refs.append((name, lref))
        off += 8 + ln
    per = [dict(pos=[], flag=[], lseq=[], cig=[], seq=[], hp=[], ps=[], qn=[]) for _ in refs]
    n = len(buf)
    mv = memoryview(buf)
    while off + 4 <= n:
        bs = struct.unpack_from("<i", buf, off)[0]
        rec = off + 4
        rid, pos, l_name, _mq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", buf, rec)
        off = rec + bs
        if rid < 0 or (contigs is not None and refs[rid][0] not in contigs):
            continue
        p = rec + 32
        qn = bytes(mv[p:p + l_name - 1]).decode()
        p += l_name
        cig = np.frombuffer(buf, "<u4", n_cig, p)
        p += 4 * n_cig
        nb = (l_seq + 1) // 2
        seq = np.frombuffer(buf, np.uint8, nb, p)
        p += nb + l_seq
        tags = _aux_ints(buf, p, off)
        if n_cig == 2 and b"CG" in tags and (int(cig[0]) & 15) == 4 and (int(cig[0]) >> 4) == l_seq and (int(cig[1]) & 15) == 3:
            cig = tags[b"CG"]            # long-CIGAR placeholder `<l_seq>S<ref_len>N`: htslib / pysam hand the reference the real one
        d = per[rid]
        d["pos"].append(pos); d["flag"].append(flag); d["lseq"].append(l_seq); d["cig"].append(cig); d["seq"].append(seq)
        d["hp"].append(tags.get(b"HP", 0)); d["ps"].append(tags.get(b"PS", 0)); d["qn"].append(qn)
    out = []
    for (name, lref), d in zip(refs, per):
        if contigs is not None and name not in contigs:
            continue
        cl = np.array([len(c) for c in d["cig"]], np.int64)
        sl = np.array([len(s) for s in d["seq"]], np.int64)
        cig_off = np.concatenate([[0], np.cumsum(cl)]).astype(np.int64)
        seq_off = np.concatenate([[0], np.cumsum(sl)]).astype(np.int64)
        cigar = np.concatenate(d["cig"]) if d["cig"] else np.zeros(0, np.uint32)
        seq4 = np.concatenate(d["seq"]) if d["seq"] else np.zeros(0, np.uint8)
        ref = fasta.get(name) if fasta else None
        if ref is None:
            ref = np.full(lref, ord("N"), np.uint8)
        out.append(ReadSet(name, ref, d["pos"], d["flag"], cig_off, cigar, seq_off, d["lseq"], seq4, d["hp"], d["ps"], d["qn"]))
    return out, text


def read_fasta(path, contigs=None):
    """-> {contig: uint8 array of the sequence bytes, case preserved}.  With a `.fai` next to an uncompressed FASTA only the
    requested contigs are read (seek + one numpy pass that drops the line ends) — what pysam.FastaFile.fetch does for the
    reference (generate_SNP_pileups.py:135-137).  Without an index (or for gzip/BGZF input) the file is scanned."""
    with open(path, "rb") as f:
        gz = f.read(2) == b"\x1f\x8b"
    fai = path + ".fai"
    if not gz and os.path.exists(fai):
        out = {}
        with open(fai) as idx, open(path, "rb") as f:
            for line in idx:
                t = line.rstrip("\n").split("\t")
                if len(t) < 5:
                    continue
                name, length, offset, linebases, linewidth = t[0], int(t[1]), int(t[2]), int(t[3]), int(t[4])
                if contigs is not None and name not in contigs:
                    continue
                if length == 0 or linebases <= 0:
                    out[name] = np.zeros(0, np.uint8)
                    continue
                nlines = (length + linebases - 1) // linebases
                f.seek(offset)
                raw = np.frombuffer(f.read((nlines - 1) * linewidth + (length - (nlines - 1) * linebases)), np.uint8)
                if linewidth == linebases or nlines == 1:
                    seq = raw[:length].copy()
                else:
                    full = (nlines - 1) * linewidth
                    body = raw[:full].reshape(nlines - 1, linewidth)[:, :linebases].reshape(-1)
                    seq = np.concatenate([body, raw[full:full + length - (nlines - 1) * linebases]])
                if len(seq) != length:
                    raise ValueError("%s: %s is shorter than its index says" % (path, name))
                out[name] = seq
        return out
    opener = gzip.open if gz else open
    out, name, chunks = {}, None, []
    with opener(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None and (contigs is None or name in contigs):
                    out[name] = np.frombuffer(b"".join(chunks), np.uint8).copy()
                name, chunks = line[1:].split()[0].decode(), []
            elif contigs is None or name in contigs:
                chunks.append(line.strip())
    if name is not None and (contigs is None or name in contigs):
        out[name] = np.frombuffer(b"".join(chunks), np.uint8).copy()
    return out


def write_fasta(path, readsets, width=60):
    with open(path, "wb") as f:
        for rs in readsets:
            f.write(b">" + rs.chrom.encode() + b"\n")
            b = rs.ref.tobytes()
            for i in range(0, len(b), width):
                f.write(b[i:i + width] + b"\n")
    with open(path + ".fai", "w") as f:
        off = 0
        for rs in readsets:
            off += len(rs.chrom) + 2
            f.write("%s\t%d\t%d\t%d\t%d\n" % (rs.chrom, rs.contig_len, off, width, width + 1))
            off += rs.contig_len + (rs.contig_len + width - 1) // width


# ---------------------------------------------------------------------------------------------- native reader (libnc_bamio.so)
import ctypes  # noqa: E402

_LIB = None


class NcBamContig(ctypes.Structure):                     # include/nanocaller_b200_io.h
    _fields_ = [("name", ctypes.c_char * 256), ("length", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("n_reads", ctypes.c_int64), ("n_cigar", ctypes.c_int64), ("n_seq", ctypes.c_int64)]


IO_EXPORTS = ["nc_bam_open", "nc_bam_open_region", "nc_bam_error", "nc_bam_n_contigs", "nc_bam_header_text", "nc_bam_contig", "nc_bam_fill",
              "nc_bam_qname", "nc_bam_write_tagged", "nc_bam_close"]


def load_io_library():
    """ctypes handle of libnc_bamio.so (built in-tree by nanocaller_b200.build)."""
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libnc_bamio.so")
        if not os.path.exists(path):
            from .. import build
            build.build_bamio()
        lib = ctypes.CDLL(path)
        vp = ctypes.c_void_p
        lib.nc_bam_open.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp)]
        lib.nc_bam_open_region.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        lib.nc_bam_error.argtypes = [vp]; lib.nc_bam_error.restype = ctypes.c_char_p
        lib.nc_bam_n_contigs.argtypes = [vp]
        lib.nc_bam_header_text.argtypes = [vp, ctypes.POINTER(ctypes.c_int64)]; lib.nc_bam_header_text.restype = ctypes.c_void_p
        lib.nc_bam_contig.argtypes = [vp, ctypes.c_int, ctypes.POINTER(NcBamContig)]
        lib.nc_bam_fill.argtypes = [vp, ctypes.c_int, ctypes.c_int] + [vp] * 9
        lib.nc_bam_qname.argtypes = [vp, ctypes.c_int, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int]
        lib.nc_bam_write_tagged.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        lib.nc_bam_close.argtypes = [vp]; lib.nc_bam_close.restype = None
        _LIB = lib
    return _LIB


def find_bai(path):
    for cand in (path + ".bai", os.path.splitext(path)[0] + ".bai"):
        if os.path.exists(cand):
            return cand
    return None


def read_bam_native(path, fasta=None, contigs=None, threads=0, alloc=None, qnames=False, use_index=True):
    """Same result as `read_bam`, through libnc_bamio.so: parallel BGZF inflate + parallel record copy into the
    staging arrays.  `alloc(shape, dtype)` lets the caller provide the arrays (e.g. pinned host memory for the H2D copy);
    query names are only materialised with `qnames=True` (the device path has no use for them).  With `contigs` given and a
    BAI index next to the file, only the blocks of those contigs are inflated (nc_bam_open_region)."""
    lib = load_io_library()
    h = ctypes.c_void_p()
    bai = find_bai(path) if (contigs is not None and use_index) else None
    if bai:
        names = [c.encode() for c in sorted(contigs)]
        arr = (ctypes.c_char_p * max(1, len(names)))(*names)
        rc = lib.nc_bam_open_region(os.fsencode(path), os.fsencode(bai), arr, len(names), int(threads), ctypes.byref(h))
    else:
        rc = lib.nc_bam_open(os.fsencode(path), int(threads), ctypes.byref(h))
    try:
        if rc != 0:
            msg = (lib.nc_bam_error(h) or b"").decode() if h else "open failed"
            if rc == -2:
                raise FileNotFoundError("%s: %s" % (path, msg))
            raise ValueError("%s: %s" % (path, msg))
        ln = ctypes.c_int64()
        tp = lib.nc_bam_header_text(h, ctypes.byref(ln))
        text = ctypes.string_at(tp, ln.value).decode(errors="replace") if tp and ln.value else ""
        mk = alloc or (lambda shape, dtype: np.empty(shape, dtype))
        out = []
        for i in range(lib.nc_bam_n_contigs(h)):
            c = NcBamContig()
            lib.nc_bam_contig(h, i, ctypes.byref(c))
            name = c.name.decode()
            if contigs is not None and name not in contigs:
                continue
            n = c.n_reads
            pos, flag, lseq = mk(n, np.int32), mk(n, np.uint16), mk(n, np.int32)
            cig_off, seq_off = mk(n + 1, np.int64), mk(n + 1, np.int64)
            cigar, seq4 = mk(c.n_cigar, np.uint32), mk(c.n_seq, np.uint8)
            hp, ps = np.empty(n, np.int8), np.empty(n, np.int32)
            rc = lib.nc_bam_fill(h, i, int(threads), *[a.ctypes.data for a in (pos, flag, cig_off, cigar, seq_off, lseq, seq4, hp, ps)])
            if rc != 0:
                raise ValueError("%s: nc_bam_fill failed (%d)" % (path, rc))
            qn = None
            if qnames:
                buf = ctypes.create_string_buffer(256)
                qn = []
                for k in range(n):
                    lib.nc_bam_qname(h, i, k, buf, 256)
                    qn.append(buf.value.decode())
            ref = fasta.get(name) if fasta else None
            if ref is None:
                ref = np.full(c.length, ord("N"), np.uint8)
            out.append(ReadSet(name, ref, pos, flag, cig_off, cigar, seq_off, lseq, seq4, hp, ps, qn))
        return out, text
    finally:
        if h:
            lib.nc_bam_close(h)


def write_haplotagged_bam(path, chrom, hp, ps, out_path, level=4, threads=0):
    """`{contig}.phased.bam` of the reference (indelCaller.py:244-245): the records of `chrom` in `path`, copied whole, with HP / PS
    replaced by the given per-read tags (file order of the contig's mapped records, i.e. the ReadSet's order).  Uses the BAI when
    there is one, so only that contig is inflated."""
    lib = load_io_library()
    h = ctypes.c_void_p()
    bai = find_bai(path)
    if bai:
        arr = (ctypes.c_char_p * 1)(chrom.encode())
        rc = lib.nc_bam_open_region(os.fsencode(path), os.fsencode(bai), arr, 1, int(threads), ctypes.byref(h))
    else:
        rc = lib.nc_bam_open(os.fsencode(path), int(threads), ctypes.byref(h))
    try:
        if rc != 0:
            raise ValueError("%s: %s" % (path, (lib.nc_bam_error(h) or b"").decode() if h else "open failed"))
        idx = -1
        c = NcBamContig()
        for i in range(lib.nc_bam_n_contigs(h)):
            lib.nc_bam_contig(h, i, ctypes.byref(c))
            if c.name.decode() == chrom:
                idx = i
                break
        if idx < 0:
            raise KeyError("contig %r is not in %s" % (chrom, path))
        hp = np.ascontiguousarray(hp, np.int8)
        ps = np.ascontiguousarray(ps, np.int32)
        if len(hp) != c.n_reads or len(ps) != c.n_reads:
            raise ValueError("%d tags for %d records of %s" % (len(hp), c.n_reads, chrom))
        rc = lib.nc_bam_write_tagged(h, idx, hp.ctypes.data_as(ctypes.c_void_p), ps.ctypes.data_as(ctypes.c_void_p), os.fsencode(out_path),
                                     int(level), int(threads))
        if rc != 0:
            raise IOError("nc_bam_write_tagged(%s) failed (%d)" % (out_path, rc))
    finally:
        if h:
            lib.nc_bam_close(h)
    return out_path


def bam_contigs(path):
    """Ordered {contig: length} from the BAM header (`sam_file.references` / `get_reference_length`, utils.py:9-50); reads only
    the head of the file."""
    with gzip.open(path, "rb") as f:
        head = f.read(12)
        if head[:4] != b"BAM\x01":
            raise ValueError("%s: not a BAM file" % path)
        l_text = struct.unpack_from("<i", head, 4)[0]
        rest = head[8:] + f.read(l_text)                     # 4 bytes already read past l_text
        n_ref = struct.unpack_from("<i", rest, l_text)[0]
        out = {}
        for _ in range(n_ref):
            ln = struct.unpack("<i", f.read(4))[0]
            name = f.read(ln)[:-1].decode()
            out[name] = struct.unpack("<i", f.read(4))[0]
        return out


def open_alignment_device(ctx, sam_path, fasta_path, contigs=None):
    """Like `open_alignment`, but the BAM is inflated and decoded on the GPU (nc_bam_device_open): registers `DeviceContig`s.
    Raises capi.NcError when the file does not fit the device budget or is not a BGZF BAM (the caller falls back to the host reader)."""
    from . import sources
    if not os.path.exists(sam_path):
        raise FileNotFoundError(sam_path)
    table = ctx.bam_device_open(sam_path)
    names = [t[0] for t in table if contigs is None or t[0] in contigs]
    fasta = read_fasta(fasta_path, set(names)) if fasta_path and os.path.exists(fasta_path) else {}
    out = []
    for i, (name, length, n_reads, n_tagged) in enumerate(table):
        if contigs is not None and name not in contigs:
            continue
        ref = fasta.get(name)
        if ref is None:
            ref = np.full(length, ord("N"), np.uint8)
        out.append(sources.DeviceContig(ctx, sam_path, i, name, ref, length, n_reads, n_tagged))
    sources._REGISTRY[sam_path] = {c.chrom: c for c in out}
    return out


def open_alignment(sam_path, fasta_path, native=True, contigs=None):
    """Parse `sam_path` (BAM) and `fasta_path` once and register the contigs (all, or the named ones) as an alignment source."""
    from . import sources
    if not os.path.exists(sam_path):
        raise FileNotFoundError(sam_path)
    fasta = read_fasta(fasta_path, contigs) if fasta_path and os.path.exists(fasta_path) else None
    readsets, _ = (read_bam_native if native else read_bam)(sam_path, fasta, contigs=contigs)
    sources.register_source(sam_path, readsets)
    return readsets
