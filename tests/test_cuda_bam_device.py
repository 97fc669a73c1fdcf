"""GPU: device-side BAM input (nc_bam_device_open / nc_bam_device_stage: BGZF inflate, record walk and record decoding on the GPU)
against the host readers on the same files, array for array — stored, fixed-Huffman and dynamic-Huffman DEFLATE blocks, several
contigs, HP / PS tags, a long-CIGAR (CG:B,I) read — and the SNP path from a file staged that way."""
import struct
import zlib

import numpy as np
import pytest

from nanocaller_b200.host import bamio
from nanocaller_b200.synth import make_world
from tests.golden.cases import _handmade

pytestmark = pytest.mark.gpu

KEYS = ("pos", "flag", "cigar_off", "cigar", "seq_off", "l_seq", "seq4", "hp", "ps")


def _rewrite(src, dst, level, strategy, block=0xff00):
    """Same BAM stream, re-deflated block by block with the given zlib level / strategy."""
    raw = bamio.bgzf_decompress(src)
    out = []
    for off in range(0, len(raw), block):
        chunk = raw[off:off + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        body = co.compress(chunk) + co.flush()
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(body) + 25) + body +
                   struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    with open(dst, "wb") as f:
        f.write(b"".join(out) + bamio._BGZF_EOF)


def _long_cigar_set():
    from nanocaller_b200.host.readset import ReadSet
    rng = np.random.default_rng(11)
    n_ops = 70_001
    ops = np.empty(n_ops, np.uint32)
    ops[0::2] = (2 << 4) | 0
    ops[1::2] = (1 << 4) | 2
    l_long = 2 * 35_001
    short = np.array([(50 << 4) | 0], np.uint32)
    contig = 3 * 35_001 + 200
    ref = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, contig)]
    lseq = np.array([50, l_long, 50], np.int32)
    seq_off = np.concatenate([[0], np.cumsum((lseq.astype(np.int64) + 1) // 2)])
    nib = np.array([1, 2, 4, 8], np.uint8)[rng.integers(0, 4, int(seq_off[-1]) * 2)]
    seq4 = (nib[0::2] << 4 | nib[1::2]).astype(np.uint8)
    return ReadSet("chrL", ref, [5, 10, 60], [0, 16, 0], np.array([0, 1, 1 + n_ops, 2 + n_ops], np.int64), np.concatenate([short, ops, short]), seq_off, lseq, seq4,
                   hp=[1, 2, 0], ps=[7, 7, 0])


@pytest.mark.parametrize("level,strategy", [(4, zlib.Z_DEFAULT_STRATEGY), (0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (9, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_HUFFMAN_ONLY)])
def test_device_reader_equals_host_reader(tmp_path, level, strategy):
    from nanocaller_b200.host import snp_pileups
    rs1 = make_world(chrom="chrA", preset="ont", contig_len=300_000, seed=5, coverage=20.0, indel_every=900, indel_maxlen=9, junk_frac=0.05, untagged_frac=0.2).reads
    rs2 = _handmade()
    rs3 = _long_cigar_set()
    rs4 = make_world(chrom="chrB", preset="hifi", contig_len=120_000, seed=6, coverage=12.0).reads
    base, bam, fa = str(tmp_path / "base.bam"), str(tmp_path / "x.bam"), str(tmp_path / "x.fa")
    bamio.write_bam(base, [rs1, rs2, rs3, rs4])
    bamio.write_fasta(fa, [rs1, rs2, rs3, rs4])
    _rewrite(base, bam, level, strategy)
    fasta = bamio.read_fasta(fa)
    want, _ = bamio.read_bam_native(bam, fasta)
    ctx = snp_pileups.context(0)
    snp_pileups._staged.clear()
    contigs = ctx.bam_device_open(bam)
    assert [(c[0], c[1], c[2]) for c in contigs] == [(w.chrom, w.contig_len, w.n) for w in want]
    assert [c[3] for c in contigs] == [int(((w.hp == 1) | (w.hp == 2)).sum()) for w in want]
    for i, w in enumerate(want):
        ctx.bam_device_stage(i, w.ref)
        got = ctx.fetch_staged()
        for k in KEYS:
            np.testing.assert_array_equal(got[k], getattr(w, k), err_msg="%s %s" % (w.chrom, k))
    tm = ctx.bam_device_timings()
    assert tm["inflated_bytes"] > tm["compressed_bytes"] * (0.9 if level == 0 else 1.0) and tm["inflate_ms"] > 0
    ctx.bam_device_close()


def test_corrupt_block_is_reported(tmp_path):
    from nanocaller_b200.host import capi, snp_pileups
    rs = make_world(chrom="chrA", preset="ont", contig_len=60_000, seed=5, coverage=10.0).reads
    bam = str(tmp_path / "c.bam")
    bamio.write_bam(bam, [rs])
    raw = bytearray(open(bam, "rb").read())
    for k in range(4000, 4200):
        raw[k] ^= 0x5A
    open(bam, "wb").write(bytes(raw))
    ctx = snp_pileups.context(0)
    with pytest.raises(capi.NcError):
        ctx.bam_device_open(bam)


def test_snp_scan_from_a_device_staged_file_matches_host_staging(tmp_path):
    """K0-K2 on a contig staged by the device reader give the same sites, tensors and metadata as on the host-staged arrays."""
    from nanocaller_b200.host import capi, snp_pileups
    rs = make_world(chrom="chr20", preset="ont", contig_len=400_000, seed=20, coverage=30.0).reads
    bam = str(tmp_path / "s.bam")
    bamio.write_bam(bam, [rs])
    dct = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont", supplementary=False)
    params = capi.snp_params(dct, "diploid")
    ctx = snp_pileups.context(0)
    snp_pileups._staged.clear()
    ctx.stage_reads(rs)
    n0 = ctx.snp_scan(params, [(1, 400_000)])
    mat0, meta0, depth0, _ = ctx.snp_fetch()
    ctx.bam_device_open(bam)
    ctx.bam_device_stage(0, rs.ref)
    n1 = ctx.snp_scan(params, [(1, 400_000)])
    mat1, meta1, depth1, _ = ctx.snp_fetch()
    assert n0 == n1 > 3000 and np.array_equal(mat0[:, :1025], mat1[:, :1025]) and meta0.tobytes() == meta1.tobytes() and depth0[0] == depth1[0]
    ctx.bam_device_close()


def test_bai_seeded_parallel_walk_equals_single_walker(tmp_path):
    """With a BAI next to the file the record chain is followed from all the record starts the index lists; same arrays as without."""
    from nanocaller_b200.host import snp_pileups
    rs1 = make_world(chrom="chrA", preset="ont", contig_len=900_000, seed=5, coverage=20.0, indel_every=900, indel_maxlen=9, untagged_frac=0.2).reads
    rs2 = make_world(chrom="chrB", preset="hifi", contig_len=500_000, seed=6, coverage=12.0).reads
    bam, plain = str(tmp_path / "i.bam"), str(tmp_path / "p.bam")
    bamio.write_bam(bam, [rs1, rs2], index=True)
    bamio.write_bam(plain, [rs1, rs2])
    ctx = snp_pileups.context(0)
    snp_pileups._staged.clear()
    got = {}
    for tag, path in (("indexed", bam), ("plain", plain)):
        table = ctx.bam_device_open(path)
        got[tag] = (table, ctx.bam_device_timings()["record_walk"])
        for i, w in enumerate((rs1, rs2)):
            ctx.bam_device_stage(i, w.ref)
            arr = ctx.fetch_staged()
            for k in KEYS:
                want = getattr(w, k) if k != "ps" else np.where(w.hp > 0, w.ps, 0)
                np.testing.assert_array_equal(arr[k], want, err_msg="%s %s %s" % (tag, w.chrom, k))
    assert got["indexed"][0] == got["plain"][0]
    assert got["indexed"][1].startswith("parallel") and got["plain"][1] == "single walker"
    ctx.bam_device_close()
