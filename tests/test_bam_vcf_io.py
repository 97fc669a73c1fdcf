"""CPU: BAM/BGZF/FASTA round trip into the staging arrays, and VCF header / sort / PASS filter / BGZF output."""
import gzip
import os

import numpy as np

from nanocaller_b200.host import bamio, sources, vcfio
from nanocaller_b200.synth import make_world
from tests.golden.cases import _handmade


def _same(a, b):
    for k in ("pos", "flag", "cigar_off", "cigar", "seq_off", "l_seq", "seq4", "hp", "ps", "ref"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=k)


def test_bam_fasta_round_trip(tmp_path):
    rs1 = make_world(chrom="chrA", preset="ont", contig_len=30_000, seed=3, coverage=8.0, indel_every=900, indel_maxlen=9,
                     junk_frac=0.05, untagged_frac=0.2).reads
    rs2 = _handmade()
    bam, fa = str(tmp_path / "x.bam"), str(tmp_path / "x.fa")
    bamio.write_bam(bam, [rs1, rs2])
    bamio.write_fasta(fa, [rs1, rs2])
    raw = open(bam, "rb").read()
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw.endswith(bamio._BGZF_EOF)
    got, text = bamio.read_bam(bam, bamio.read_fasta(fa))
    assert [g.chrom for g in got] == ["chrA", "tiny"] and "@SQ\tSN:chrA" in text
    # the handmade set has no tags for junk twins etc.; ps is only defined where hp > 0
    for a, b in zip(got, [rs1, rs2]):
        b2 = b
        b2.ps = np.where(b.hp > 0, b.ps, 0).astype(np.int32)
        _same(a, b2)
    # path-based resolution used by the drop-in functions
    sources.unregister_all()
    sources.attach_fasta(bam, fa)
    r = sources.resolve(bam, "tiny")
    assert r.n == rs2.n and r.checksum() == got[1].checksum()


def test_vcf_writer(tmp_path):
    lines = ["chr2\t50\t.\tA\tG\t10.000\tPASS\tPR=0.1000,0.1000,0.9000,0.1000;FQ=0.5000\tGT:DP:VF:AD:ADF:ADR\t1/1:9:0.5:1,2:1,1:0,1\n",
             "chr1\t70\t.\tC\t.\t3.000\tREF\tPR=0.1000,0.9000,0.1000,0.1000;FQ=0.2000\tGT:DP:VF:AD:ADF:ADR\t./.:9:.:.:.:.\n",
             "chr1\t20\t.\tT\tA\t30.000\tPASS\tPR=0.9000,0.1000,0.1000,0.6000;FQ=0.4000\tGT:DP:VF:AD:ADF:ADR\t0/1:9:0.4:5,4:2,2:3,2\n",
             "chr1\t20\t.\tT\tA\t31.000\tPASS\tPR=0.9000,0.1000,0.1000,0.6000;FQ=0.4000\tGT:DP:VF:AD:ADF:ADR\t0/1:9:0.4:5,4:2,2:3,2\n"]
    s = vcfio.sort_records(lines, ["chr1", "chr2"])
    assert [x.split("\t")[:2] for x in s] == [["chr1", "20"], ["chr1", "20"], ["chr1", "70"], ["chr2", "50"]]
    assert s[0].split("\t")[5] == "30.000"                           # stable: boundary duplicates keep their order
    assert len(vcfio.pass_only(s)) == 3
    p = str(tmp_path / "o.vcf.gz")
    vcfio.write_vcf(p, "snps", ["chr1", "chr2"], lines, sample="S1")
    txt = gzip.open(p, "rt").read()
    assert txt.startswith("##fileformat=VCFv4.2\n") and "##contig=<ID=chr1>\n##contig=<ID=chr2>\n" in txt
    assert txt.splitlines()[-5].endswith("FORMAT\tS1") and txt.endswith(s[-1])
    h = vcfio.header("indels", ["chr1"])
    assert "ID=GQ" in h and "ID=PS" in h and "LOW" not in h


def test_native_bam_reader_matches_python_reader(tmp_path):
    """libnc_bamio.so (parallel inflate + record copy) vs the pure-Python parser on the same files, array for array."""
    rs1 = make_world(chrom="chrA", preset="ont", contig_len=60_000, seed=5, coverage=12.0, indel_every=900, indel_maxlen=9,
                     junk_frac=0.05, untagged_frac=0.2).reads
    rs2 = _handmade()
    rs3 = make_world(chrom="chrB", preset="ont", contig_len=20_000, seed=6, coverage=5.0).reads
    bam, fa = str(tmp_path / "n.bam"), str(tmp_path / "n.fa")
    bamio.write_bam(bam, [rs1, rs2, rs3])
    bamio.write_fasta(fa, [rs1, rs2, rs3])
    fasta = bamio.read_fasta(fa)
    want, text_w = bamio.read_bam(bam, fasta)
    for threads in (1, 4):
        got, text_g = bamio.read_bam_native(bam, fasta, threads=threads, qnames=True)
        assert text_g == text_w and [g.chrom for g in got] == [w.chrom for w in want]
        for g, w in zip(got, want):
            _same(g, w)
            assert g.qnames == w.qnames and g.contig_len == w.contig_len
    only, _ = bamio.read_bam_native(bam, fasta, contigs={"tiny"})
    assert [r.chrom for r in only] == ["tiny"] and only[0].checksum() == want[1].checksum()
    # caller-provided buffers (what a pinned-memory stager passes)
    pool = []

    def alloc(shape, dtype):
        a = np.zeros(shape, dtype)
        pool.append(a)
        return a
    got, _ = bamio.read_bam_native(bam, fasta, alloc=alloc)
    assert len(pool) == 7 * 3 and got[0].pos is pool[0] and got[0].checksum() == want[0].checksum()


def test_native_bam_reader_errors(tmp_path):
    import pytest
    with pytest.raises(FileNotFoundError):
        bamio.read_bam_native(str(tmp_path / "absent.bam"))
    p = str(tmp_path / "garbage.bam")
    open(p, "wb").write(b"this is not a BGZF stream at all, just text......")
    with pytest.raises(ValueError, match="BGZF"):
        bamio.read_bam_native(p)
    p2 = str(tmp_path / "notbam.bam")
    open(p2, "wb").write(bamio.bgzf_compress(b"VCF\x01 something else"))
    with pytest.raises(ValueError, match="magic"):
        bamio.read_bam_native(p2)
    # truncated file: cut the valid BAM in the middle of a block
    rs = _handmade()
    p3 = str(tmp_path / "ok.bam")
    bamio.write_bam(p3, [rs])
    raw = open(p3, "rb").read()
    p4 = str(tmp_path / "cut.bam")
    open(p4, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(ValueError):
        bamio.read_bam_native(p4)


def test_io_header_symbols_exported():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "nanocaller_b200_io.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(nc_bam_[a-z0-9_]+)\s*\(", src)))
    lib = bamio.load_io_library()
    assert names == sorted(bamio.IO_EXPORTS) and all(hasattr(lib, n) for n in names)


def test_write_vcf_blobs_equals_line_writer(tmp_path):
    """The blob writer used by the command line (records formatted by the library) against sort_records / pass_only / write_vcf on the
    same records: several parts per contig, out-of-order regions, shared-boundary duplicates, PASS filter, threaded BGZF."""
    from nanocaller_b200.host import capi
    rng = np.random.RandomState(3)
    parts, lines = [], []
    for chrom, lo, hi, n in (("chr2", 1, 9000, 300), ("chr1", 5000, 9000, 200), ("chr1", 1, 5000, 250), ("chr3", 1, 100, 0)):
        pos = np.sort(rng.randint(lo, hi + 1, n)).astype(np.int32)
        if n:
            pos[-1] = hi
            pos[0] = lo                                            # chr1: 5000 appears in two parts (shared boundary)
        probs = rng.rand(n, 4).astype(np.float32)
        ref = rng.randint(0, 4, n).astype(np.uint8)
        fwd = rng.randint(0, 40, (n, 4)).astype(np.uint16); rev = rng.randint(0, 40, (n, 4)).astype(np.uint16)
        dp = (fwd.sum(1) + rev.sum(1) + 1).astype(np.int32)
        alt = np.minimum(dp, rng.randint(0, 30, n)).astype(np.int32)
        blob, off, ok = capi.format_snp_records(chrom, pos, ref, probs, dp, alt, fwd, rev)
        parts.append((chrom, blob, off, ok, pos))
        lines += [blob[off[i]:off[i + 1]].decode() for i in range(n) if off[i + 1] > off[i]]
    contigs = ["chr1", "chr2", "chr3"]
    for pass_only in (False, True):
        a, b = str(tmp_path / ("a%d.vcf.gz" % pass_only)), str(tmp_path / ("b%d.vcf.gz" % pass_only))
        sel = vcfio.pass_only(lines) if pass_only else lines
        vcfio.write_vcf(a, "snps", contigs, sel, "S")
        n = vcfio.write_vcf_blobs(b, "snps", contigs, parts, "S", pass_only=pass_only)
        ta, tb = gzip.open(a, "rt").read(), gzip.open(b, "rt").read()
        assert ta == tb and n == len(sel) > 100
        c = str(tmp_path / ("c%d.vcf.gz" % pass_only))
        vcfio.write_vcf_blobs(c, "snps", contigs, parts, "S", pass_only=pass_only, index=True)      # the vectorised CSI path
        assert gzip.open(c, "rt").read() == ta
        for chrom, beg, end in (("chr1", 0, 10_000), ("chr1", 4990, 5010), ("chr2", 100, 2000), ("chr3", 0, 50), ("chr1", 8999, 9000)):
            want = [ln for ln in sel if ln.split("\t")[0] == chrom and beg <= int(ln.split("\t")[1]) - 1 < end]
            want = [ln for ln in vcfio.sort_records(want, contigs)]
            assert vcfio.csi_query(c, chrom, beg, end) == want, (chrom, beg, end)
    big = os.urandom(5_000_000)
    z = bamio.bgzf_compress(big)
    p = str(tmp_path / "big.gz")
    open(p, "wb").write(z)
    assert gzip.open(p, "rb").read() == big and z.endswith(bamio._BGZF_EOF)


def test_fasta_index_reader_equals_scan(tmp_path):
    """read_fasta through the .fai (seek + numpy, selected contigs only) gives the bytes of the line-by-line scan: line widths that do
    and do not divide the contig length, a one-line contig, lower-case soft-masked bases preserved."""
    rng = np.random.RandomState(8)
    from nanocaller_b200.host.readset import ReadSet
    sets = []
    for name, n in (("a", 1234), ("b", 60), ("c", 61), ("d", 7), ("e", 120)):
        ref = rng.choice(np.frombuffer(b"ACGTacgtN", np.uint8), n)
        sets.append(ReadSet(name, ref, [], [], [0], np.zeros(0, np.uint32), [0], [], np.zeros(0, np.uint8)))
    fa = str(tmp_path / "r.fa")
    bamio.write_fasta(fa, sets, width=60)
    via_index = bamio.read_fasta(fa)
    os.rename(fa + ".fai", fa + ".fai.off")
    via_scan = bamio.read_fasta(fa)
    os.rename(fa + ".fai.off", fa + ".fai")
    assert list(via_index) == list(via_scan) == [s.chrom for s in sets]
    for s_ in sets:
        np.testing.assert_array_equal(via_index[s_.chrom], s_.ref)
        np.testing.assert_array_equal(via_scan[s_.chrom], s_.ref)
    only = bamio.read_fasta(fa, contigs={"c", "e"})
    assert list(only) == ["c", "e"] and only["c"].tobytes() == sets[2].ref.tobytes()


def test_bam_contigs_from_header(tmp_path):
    rs1 = make_world(chrom="chrA", preset="ont", contig_len=20_000, seed=3, coverage=3.0).reads
    rs2 = _handmade()
    bam = str(tmp_path / "h.bam")
    bamio.write_bam(bam, [rs1, rs2])
    assert bamio.bam_contigs(bam) == {"chrA": 20_000, "tiny": rs2.contig_len}
    assert list(bamio.bam_contigs(bam)) == ["chrA", "tiny"]


def test_region_read_through_bai_equals_whole_file_read(tmp_path):
    """nc_bam_open_region (only the named contigs' BGZF blocks are inflated, located through the BAI) against the whole-file reader,
    for every subset of contigs, with records that straddle BGZF block boundaries and a contig without reads."""
    from nanocaller_b200.host.readset import ReadSet
    rs1 = make_world(chrom="chrA", preset="ont", contig_len=80_000, seed=5, coverage=10.0, indel_every=900, indel_maxlen=9).reads
    rs2 = _handmade()
    empty = ReadSet("void", np.frombuffer(b"ACGT" * 10, np.uint8), [], [], [0], np.zeros(0, np.uint32), [0], [], np.zeros(0, np.uint8))
    rs3 = make_world(chrom="chrB", preset="ont", contig_len=50_000, seed=6, coverage=8.0).reads
    bam = str(tmp_path / "i.bam")
    bamio.write_bam(bam, [rs1, rs2, empty, rs3], index=True)
    assert os.path.getsize(bam) > 3 * 65536 and bamio.find_bai(bam) == bam + ".bai"
    whole, text = bamio.read_bam_native(bam, use_index=False)
    by_name = {r.chrom: r for r in whole}
    for subset in ({"chrA"}, {"tiny"}, {"chrB"}, {"void"}, {"chrB", "tiny"}, {"chrA", "chrB", "tiny", "void"}, {"nope"}):
        got, text_g = bamio.read_bam_native(bam, contigs=subset, threads=3)
        assert text_g == text and sorted(g.chrom for g in got) == sorted(subset & set(by_name))
        for g in got:
            _same(g, by_name[g.chrom])
            assert g.n == by_name[g.chrom].n
    # a BAI that does not belong to the file is refused, not silently misread
    other = str(tmp_path / "o.bam")
    bamio.write_bam(other, [rs2], index=True)
    os.replace(other + ".bai", bam + ".bai")
    import pytest
    with pytest.raises(ValueError, match="BAI"):
        bamio.read_bam_native(bam, contigs={"chrA"})


def test_csi_index_round_trip(tmp_path):
    """write_indexed: BGZF VCF + CSI index (what `tabix -fp vcf --csi` leaves, snpCaller.py:283-285).  Region queries through the index
    return exactly the records a linear scan finds, for SNP records and for indel records with long REF alleles spanning bin borders."""
    import struct
    rng = np.random.RandomState(4)
    lines = []
    for chrom, n in (("chr1", 6000), ("chr2", 2500)):
        pos = np.unique(rng.randint(1, 3_000_000, n))
        for p_ in pos:
            if rng.rand() < 0.1:
                ref = "A" + "".join(rng.choice(list("ACGT"), rng.randint(1, 60)))          # deletion: a long REF allele
                lines.append("%s\t%d\t.\t%s\tA\t30.00\tPASS\t.\tGT:GQ\t1/1:30.00\n" % (chrom, p_, ref))
            else:
                lines.append("%s\t%d\t.\tC\tT\t12.000\tPASS\tPR=0.1,0.2,0.3,0.9;FQ=0.5\tGT:DP:VF:AD:ADF:ADR\t0/1:30:0.5:15,15:8,7:7,8\n" % (chrom, p_))
    # a record sitting exactly on a 16 kb window border, with a REF reaching across it
    lines.append("chr1\t%d\t.\t%s\tG\t9.00\tPASS\t.\tGT:GQ\t0|1:9.00\n" % (16384 * 7 - 2, "G" + "T" * 9))
    contigs = ["chr1", "chr2", "chr3"]
    data = (vcfio.header("all", contigs, "S") + "".join(vcfio.sort_records(lines, contigs))).encode()
    path = str(tmp_path / "x.vcf.gz")
    vcfio.write_indexed(path, data)
    assert gzip.open(path, "rb").read() == data and len(data) > 8 * 0xff00          # several BGZF blocks: virtual offsets matter
    idx = gzip.open(path + ".csi", "rb").read()
    assert idx[:4] == b"CSI\x01" and struct.unpack_from("<ii", idx, 4) == (14, 5)
    recs = [ln for ln in data.decode().splitlines(True) if not ln.startswith("#")]

    def scan(chrom, beg, end):
        out = []
        for ln in recs:
            f = ln.split("\t", 4)
            b = int(f[1]) - 1
            if f[0] == chrom and b < end and b + len(f[3]) > beg:
                out.append(ln)
        return out
    regions = [("chr1", 0, 3_000_000), ("chr2", 1_000_000, 1_000_500), ("chr1", 16384 * 7 - 5, 16384 * 7 + 1), ("chr1", 16384 * 7, 16384 * 7 + 3),
               ("chr3", 0, 1000), ("chrZ", 0, 10), ("chr2", 2_999_000, 3_100_000)]
    regions += [("chr1", int(a), int(a) + int(rng.randint(1, 200_000))) for a in rng.randint(0, 3_000_000, 25)]
    for chrom, beg, end in regions:
        assert vcfio.csi_query(path, chrom, beg, end) == scan(chrom, beg, end), (chrom, beg, end)
    # bin arithmetic against the values of the specification's own examples
    assert vcfio._reg2bin(0, 1) == 4681 and vcfio._reg2bin(0, 16385) == 585 and vcfio._reg2bin(16384, 32768) == 4682
    assert vcfio._bin_first_window(4681) == 0 and vcfio._bin_first_window(585) == 0 and vcfio._bin_first_window(586) == 8 and vcfio._bin_first_window(1) == 0


def test_haplotagged_bam_copy(tmp_path):
    """nc_bam_write_tagged: the records of one contig copied whole with new HP / PS tags — read back by the native reader and by the
    pure-Python parser; existing tags are replaced, untagged reads carry none, names survive."""
    import numpy as np
    from nanocaller_b200.host import bamio
    from nanocaller_b200.synth import make_world
    a = make_world(chrom="chrA", preset="ont", contig_len=60_000, seed=31, coverage=12.0, untagged_frac=0.3).reads
    b = make_world(chrom="chrB", preset="ont", contig_len=30_000, seed=32, coverage=8.0).reads
    src = str(tmp_path / "s.bam")
    bamio.write_bam(src, [a, b], index=True)
    rng = np.random.RandomState(3)
    hp = rng.randint(0, 3, a.n).astype(np.int8)
    ps = np.where(hp > 0, rng.randint(1, 60_000, a.n), 0).astype(np.int32)
    assert (a.hp != hp).any()
    for use_bai in (True, False):
        if not use_bai:
            import os
            os.remove(src + ".bai")
        out = bamio.write_haplotagged_bam(src, "chrA", hp, ps, str(tmp_path / ("chrA.phased.%d.bam" % use_bai)), threads=3)
        for reader in (bamio.read_bam_native, bamio.read_bam):
            got = {r.chrom: r for r in reader(out)[0]}
            assert "chrB" not in got or got["chrB"].n == 0                # header keeps both references, records are chrA's only
            g = got["chrA"]
            for f in ("pos", "flag", "cigar_off", "cigar", "seq_off", "l_seq", "seq4"):
                assert np.array_equal(getattr(g, f), getattr(a, f)), f
            assert np.array_equal(g.hp, hp) and np.array_equal(g.ps[hp > 0], ps[hp > 0]) and not g.ps[hp == 0].any()
        names = [r for r in bamio.read_bam_native(out, qnames=True)[0] if r.chrom == "chrA"][0]
        assert [names.qname(k) for k in (0, 1, a.n - 1)] == [a.qname(k) for k in (0, 1, a.n - 1)]
    import pytest
    with pytest.raises(ValueError):
        bamio.write_haplotagged_bam(src, "chrA", hp[:-1], ps[:-1], str(tmp_path / "x.bam"))


def test_haplotagged_bam_copy_in_batches(tmp_path):
    """More than one 64 MB batch of uncompressed records: the writer deflates and writes as it goes; the copy reads back identical."""
    import numpy as np
    from nanocaller_b200.host import bamio
    from nanocaller_b200.synth import make_world
    a = make_world(chrom="chrL", preset="ont", contig_len=3_000_000, seed=33, coverage=30.0, untagged_frac=1.0).reads
    src = str(tmp_path / "l.bam")
    bamio.write_bam(src, [a])
    approx = int((a.l_seq.astype(np.int64) * 3 // 2 + 4 * np.diff(a.cigar_off) + 40).sum())
    assert approx > 2.2 * 0xff00 * 1024                          # at least three batches
    hp = (np.arange(a.n) % 3).astype(np.int8)
    ps = np.where(hp > 0, 1234, 0).astype(np.int32)
    out = bamio.write_haplotagged_bam(src, "chrL", hp, ps, str(tmp_path / "l.phased.bam"))
    g = bamio.read_bam_native(out)[0][0]
    for f in ("pos", "flag", "cigar_off", "cigar", "seq_off", "l_seq", "seq4"):
        assert np.array_equal(getattr(g, f), getattr(a, f)), f
    assert np.array_equal(g.hp, hp) and np.array_equal(g.ps, ps)


def test_vcf_headers_equal_the_reference_literals():
    """tests/golden/reference_vcf_headers.json = the header lines the reference writes (snpCaller.py:259-276, indelCaller.py:373-383,
    extracted by tests/golden/make_cli_flags.py; `%s` stands for the contig / sample name)."""
    import json
    import os
    from nanocaller_b200.host import vcfio
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vcf_headers.json")))
    for kind in ("snps", "indels"):
        ref = [ln.replace("%s", "X") for ln in want[kind]]
        got = [ln for ln in vcfio.header(kind, ["X"], "X").split("\n") if ln]
        assert got == ref and len(ref) >= 7, kind


def test_long_cigar_cg_tag_round_trip(tmp_path):
    """A read with more than 65,535 CIGAR operations is stored as `<l_seq>S<ref_len>N` + CG:B,I (SAM spec 4.2.2; routine for the
    ultra-long reads of the ul_ont presets).  htslib restores the real CIGAR for the reference; both readers here must too."""
    from nanocaller_b200.host.readset import ReadSet
    rng = np.random.default_rng(11)
    n_ops = 70_001                                         # alternating 2M 1D ... : 35,001 M + 35,000 D
    ops = np.empty(n_ops, np.uint32)
    ops[0::2] = (2 << 4) | 0
    ops[1::2] = (1 << 4) | 2
    l_long = 2 * 35_001
    short = np.array([(50 << 4) | 0], np.uint32)
    contig = 3 * 35_001 + 200
    ref = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, contig)]
    lseq = np.array([50, l_long, 50], np.int32)
    seq_bytes = (lseq.astype(np.int64) + 1) // 2
    seq_off = np.concatenate([[0], np.cumsum(seq_bytes)])
    nib = np.array([1, 2, 4, 8], np.uint8)[rng.integers(0, 4, int(seq_off[-1]) * 2)]
    seq4 = (nib[0::2] << 4 | nib[1::2]).astype(np.uint8)
    cigar = np.concatenate([short, ops, short])
    cig_off = np.array([0, 1, 1 + n_ops, 2 + n_ops], np.int64)
    rs = ReadSet("chrL", ref, [5, 10, 60], [0, 16, 0], cig_off, cigar, seq_off, lseq, seq4, hp=[1, 2, 0], ps=[7, 7, 0])
    bam, fa = str(tmp_path / "l.bam"), str(tmp_path / "l.fa")
    bamio.write_bam(bam, [rs], index=True)
    bamio.write_fasta(fa, [rs])
    fasta = bamio.read_fasta(fa)
    # the file really holds the placeholder: n_cigar_op of the long read is 2
    import struct
    buf = bamio.bgzf_decompress(bam)
    assert buf.count(b"CGBI" + struct.pack("<i", n_ops)) == 1
    got_py, _ = bamio.read_bam(bam, fasta)
    _same(got_py[0], rs)
    for threads in (1, 3):
        got_n, _ = bamio.read_bam_native(bam, fasta, threads=threads)
        _same(got_n[0], rs)
    assert int(got_py[0].ref_end[1]) == 10 + 3 * 35_000 + 2
