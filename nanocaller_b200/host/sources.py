"""Alignment sources: what `dct['sam_path']` / `dct['fasta_path']` resolve to on the host.

The reference opens the BAM and the FASTA by path inside every call (generate_SNP_pileups.py:134-135).
Here a path resolves to per-contig `ReadSet`s in BAM-native encoding.  In-memory sources (the synthetic
generator, tests) are registered under a `mem://` name; file-backed BAM/FASTA sources plug in through
the same `resolve` call."""

_REGISTRY = {}
_BEDS = {}


def register_source(path, readsets):
    """Expose `readsets` (ReadSet or list of ReadSet, one per contig) under `path`."""
    if not isinstance(readsets, (list, tuple)):
        readsets = [readsets]
    _REGISTRY[path] = {rs.chrom: rs for rs in readsets}


class _Lazy:
    """A contig whose ReadSet is made on first use (a rank of a multi-GPU run only materialises the contigs it owns)."""

    def __init__(self, length, factory):
        self.length, self.factory = int(length), factory


def register_lazy(path, contigs):
    """contigs: ordered {chrom: (length, factory)}; factory() -> ReadSet, called on the first `resolve` of that contig."""
    _REGISTRY[path] = {c: _Lazy(n, f) for c, (n, f) in contigs.items()}


def contig_lengths(path):
    """Ordered {contig: length} of a registered in-memory source (what the BAM header gives for a file, utils.py:9-50)."""
    return {c: (v.length if isinstance(v, _Lazy) else v.contig_len) for c, v in _REGISTRY[path].items()}


def release(path, chrom):
    """Drop a materialised lazy contig (frees host memory once its stage is done)."""
    _REGISTRY.get(path, {}).pop(chrom, None)


class DeviceContig:
    """A contig whose reads were inflated and decoded on the GPU (Context.bam_device_open): the host only holds the reference
    bytes; staging is a device-to-device step (Context.bam_device_stage).  `materialise()` swaps in a host ReadSet (native
    reader) for the steps that need the reads on the host: phasing of untagged reads, chunk-window sharding."""

    def __init__(self, ctx, path, index, chrom, ref, length, n_reads, n_tagged):
        self.ctx, self.path, self.index, self.chrom, self.ref = ctx, path, index, chrom, ref
        self.contig_len, self.n, self.n_tagged = int(length), int(n_reads), int(n_tagged)

    def materialise(self):
        from . import bamio
        fasta = {self.chrom: self.ref}
        sets, _ = bamio.read_bam_native(self.path, fasta, contigs={self.chrom})
        rs = sets[0]
        _REGISTRY[self.path][self.chrom] = rs
        return rs


def host_reads(sam_path, chrom):
    """`resolve`, but always a host ReadSet."""
    rs = resolve(sam_path, chrom)
    return rs.materialise() if isinstance(rs, DeviceContig) else rs


def register_bed(path, intervals):
    """intervals: {chrom: [(start, end), ...]} — stands for a tabix-indexed exclude BED."""
    _BEDS[path] = intervals


def unregister_all():
    _REGISTRY.clear()
    _BEDS.clear()


_FASTA_FOR = {}


def attach_fasta(sam_path, fasta_path):
    """Remember which FASTA goes with a BAM path (the reference passes both in `dct`)."""
    _FASTA_FOR[sam_path] = fasta_path


def resolve(sam_path, chrom):
    if sam_path not in _REGISTRY:
        import os
        if os.path.exists(sam_path):               # a real BAM on disk: parse once, keep per-contig arrays
            from . import bamio
            bamio.open_alignment(sam_path, _FASTA_FOR.get(sam_path))
    try:
        rs = _REGISTRY[sam_path][chrom]
    except KeyError:
        raise FileNotFoundError("no alignment source registered for %r contig %r" % (sam_path, chrom))
    if isinstance(rs, _Lazy):
        rs = _REGISTRY[sam_path][chrom] = rs.factory()
    return rs


def restrict(sam_path, windows):
    """windows: {contig: (lo0, hi0)} — replace the registered read sets of those contigs by `ReadSet.window(lo0, hi0)`."""
    reg = _REGISTRY[sam_path]
    for chrom, (lo0, hi0) in windows.items():
        if chrom in reg:
            reg[chrom] = host_reads(sam_path, chrom).window(lo0, hi0)


def contigs(sam_path):
    return list(_REGISTRY[sam_path])


def bed_intervals(path, chrom):
    """-> list of (start, end) for `chrom`, or None when the contig is absent from the BED (the reference
    then disables exclusion: generate_SNP_pileups.py:121-123)."""
    if not path:
        return None
    table = _BEDS.get(path)
    if table is None:
        raise FileNotFoundError("no exclude BED registered for %r" % path)
    ivs = table.get(chrom)
    if ivs is None:
        return None
    # IntervalTree rejects null intervals with ValueError, which the reference catches and treats as "no exclusion"
    if any(e <= s for s, e in ivs):
        return None
    return list(ivs)
