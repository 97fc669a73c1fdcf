"""ORACLE / TEST INFRASTRUCTURE ONLY — a stand-in `pysam` so the reference modules
(`/root/reference/nanocaller_src/generate_*_pileups.py`, `utils.py`) can be imported and run
UNCHANGED in the build container, where neither pysam nor htslib exists.

It serves in-memory `ReadSet`s registered under fake paths and implements exactly the slice of the
pysam API the reference touches, with htslib's pileup-column semantics
restated from sam.c `resolve_cigar2` / pysam `get_query_sequences`:

  * admission: mapped and `flag & flag_filter == 0`; columns ascending, only where >=1 read covers,
    clipped to [start, end) when truncate=True; reads inside a column in BAM (file) order;
  * per read token: base letter (lower-case on the reverse strand), '*' inside a deletion,
    '>'/'<' inside a reference skip; with add_indels, at the LAST reference position of an op:
    next op I -> '+L<bases>' (consecutive I summed, P skipped); next op D (current not D) ->
    '-L' + 'N'*L (consecutive D summed);
  * `query_position_or_next`: query index of the base, or of the next base inside a deletion.

htslib itself is not available here, so this contract is "parity unpinned" against real pysam;
it is the specification both the numpy oracle and the CUDA kernels are held to.
Never imported by the product package.
"""
import numpy as np

_REGISTRY = {}   # path -> {"contigs": {chrom: ReadSet}, "order": [chrom,...]}
_BEDS = {}       # path -> {chrom: [(start, end), ...]}

_REF_CONSUME = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1] + [0] * 7, dtype=np.int64)
_QRY_CONSUME = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1] + [0] * 7, dtype=np.int64)
_NIB_UP = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)
_NIB_LO = np.frombuffer(b"=acmgrsvtwyhkdbn", dtype=np.uint8)


def register(path, readsets):
    """Expose `readsets` (ReadSet or list of ReadSet, one per contig) under the fake file `path`
    (used for both the BAM and the FASTA)."""
    if not isinstance(readsets, (list, tuple)):
        readsets = [readsets]
    _REGISTRY[path] = {"contigs": {rs.chrom: rs for rs in readsets}, "order": [rs.chrom for rs in readsets]}


def register_bed(path, intervals):
    """intervals: {chrom: [(start, end), ...]} served by TabixFile(path)."""
    _BEDS[path] = intervals


def unregister_all():
    _REGISTRY.clear()
    _BEDS.clear()


class AlignedSegment:
    def __init__(self, rs, i):
        self._rs, self._i = rs, i
        self.flag = int(rs.flag[i])
        self.qname = self.query_name = rs.qname(i)
        self.reference_start = self.pos = int(rs.pos[i])
        self.reference_end = int(rs.ref_end[i])
        self._seq = None

    @property
    def query_sequence(self):
        if self._seq is None:
            self._seq = self._rs.query_sequence(self._i)
        return self._seq

    seq = query_sequence

    def has_tag(self, tag):
        if tag == "HP":
            return self._rs.hp[self._i] > 0
        if tag == "PS":
            return self._rs.hp[self._i] > 0 and self._rs.ps[self._i] != 0
        return False

    def get_tag(self, tag):
        if tag == "HP" and self._rs.hp[self._i] > 0:
            return int(self._rs.hp[self._i])
        if tag == "PS" and self._rs.hp[self._i] > 0:
            return int(self._rs.ps[self._i])
        raise KeyError(tag)


class PileupRead:
    __slots__ = ("alignment", "query_position_or_next", "query_position", "is_del", "is_refskip", "indel")

    def __init__(self, alignment, qpn, is_del, is_refskip, indel):
        self.alignment = alignment
        self.query_position_or_next = qpn
        self.is_del = is_del
        self.is_refskip = is_refskip
        self.query_position = None if (is_del or is_refskip) else qpn
        self.indel = indel


class PileupColumn:
    def __init__(self, engine, pos, lo, hi):
        self._e, self.pos, self._lo, self._hi = engine, pos, lo, hi
        self.reference_pos = pos

    def get_num_aligned(self):
        return self._hi - self._lo

    @property
    def nsegments(self):
        return self._hi - self._lo

    n = nsegments

    def get_query_names(self):
        rs, ridx = self._e.rs, self._e.ridx
        return [rs.qname(int(ridx[k])) for k in range(self._lo, self._hi)]

    def get_query_sequences(self, mark_matches=False, mark_ends=False, add_indels=False):
        e = self._e
        out = []
        for k in range(self._lo, self._hi):
            tok = chr(e.ch[k])
            ind = int(e.indel[k])
            if add_indels and ind != 0:
                if ind > 0:
                    al = e.segment(int(e.ridx[k]))
                    q = int(e.qnext[k])
                    ins = al.query_sequence[q:q + ind]
                    if al.flag & 0x10:
                        ins = ins.lower()
                    tok += "+%d%s" % (ind, ins)
                else:
                    tok += "-%d%s" % (-ind, ("n" if (e.rs.flag[int(e.ridx[k])] & 0x10) else "N") * (-ind))
            out.append(tok)
        return out

    @property
    def pileups(self):
        e = self._e
        return [PileupRead(e.segment(int(e.ridx[k])), int(e.qpn[k]), bool(e.isdel[k]), bool(e.isskip[k]), int(e.indel[k]))
                for k in range(self._lo, self._hi)]


class _PileupEngine:
    """Expands all admitted reads overlapping [lo, hi) into per-(position, read) entries sorted by
    (position, BAM order)."""

    def __init__(self, rs, lo, hi, flag_filter):
        self.rs = rs
        self._segs = {}
        n = rs.n
        ref_end = rs.ref_end
        cand = np.nonzero((rs.pos < hi) & (ref_end > lo) & ((rs.flag & flag_filter) == 0) & ((rs.flag & 0x4) == 0))[0]
        P, R, CH, QPN, QNX, DEL, SKIP, IND = [], [], [], [], [], [], [], []
        for i in cand:
            cg = rs.read_cigar(i)
            ops = (cg & 15).astype(np.int64)
            lens = (cg >> 4).astype(np.int64)
            rl = lens * _REF_CONSUME[ops]
            ql = lens * _QRY_CONSUME[ops]
            rstart = int(rs.pos[i]) + np.concatenate([[0], np.cumsum(rl)[:-1]])
            qstart = np.concatenate([[0], np.cumsum(ql)[:-1]])
            nz = np.nonzero(rl > 0)[0]
            if len(nz) == 0:
                continue
            # indel annotation at the last reference position of each reference-consuming op
            op_indel = np.zeros(len(ops), np.int64)
            op_qnext = np.zeros(len(ops), np.int64)
            nc = len(ops)
            for k in nz:
                if k + 1 >= nc:
                    continue
                op, op2 = ops[k], ops[k + 1]
                if op2 == 2 and op != 2:
                    tot = lens[k + 1]
                    j = k + 2
                    while j < nc:
                        if ops[j] == 2:
                            tot += lens[j]
                        elif ops[j] in (1, 4, 0, 7, 8):
                            break
                        j += 1
                    op_indel[k] = -tot
                elif op2 == 1:
                    tot = lens[k + 1]
                    j = k + 2
                    while j < nc:
                        if ops[j] == 1:
                            tot += lens[j]
                        elif ops[j] != 6:
                            break
                        j += 1
                    op_indel[k] = tot
                    op_qnext[k] = qstart[k + 1]
                elif op2 == 6 and k + 2 < nc:
                    tot, j, qn = 0, k + 2, None
                    while j < nc:
                        if ops[j] == 1:
                            if qn is None:
                                qn = qstart[j]
                            tot += lens[j]
                        elif ops[j] != 6:
                            break
                        j += 1
                    if tot:
                        op_indel[k] = tot
                        op_qnext[k] = qn
            nib = rs.read_nibbles(i)
            lut = _NIB_LO if (rs.flag[i] & 0x10) else _NIB_UP
            opidx = np.repeat(nz, rl[nz])
            off = np.arange(len(opidx)) - np.repeat(np.cumsum(rl[nz]) - rl[nz], rl[nz])
            p = rstart[opidx] + off
            o = ops[opidx]
            is_m = (o == 0) | (o == 7) | (o == 8)
            is_d = o == 2
            is_n = o == 3
            q = np.where(is_m, qstart[opidx] + off, qstart[opidx])
            ch = np.full(len(p), ord("*"), np.uint8)
            qm = np.clip(q[is_m], 0, max(0, len(nib) - 1))
            if len(nib):
                bm = lut[nib[qm]]
                bm = np.where(q[is_m] < len(nib), bm, ord("N") if not (rs.flag[i] & 0x10) else ord("n"))
                ch[is_m] = bm
            else:
                ch[is_m] = ord("N")
            ch[is_n] = ord("<") if (rs.flag[i] & 0x10) else ord(">")
            last = off == (rl[opidx] - 1)
            ind = np.where(last, op_indel[opidx], 0)
            qnx = np.where(last, op_qnext[opidx], 0)
            keep = (p >= lo) & (p < hi)
            P.append(p[keep]); R.append(np.full(int(keep.sum()), i, np.int64)); CH.append(ch[keep])
            QPN.append(q[keep]); QNX.append(qnx[keep]); DEL.append(is_d[keep]); SKIP.append(is_n[keep]); IND.append(ind[keep])
        if P:
            P = np.concatenate(P); order = np.argsort(P, kind="stable")
            self.p = P[order]
            self.ridx = np.concatenate(R)[order]
            self.ch = np.concatenate(CH)[order]
            self.qpn = np.concatenate(QPN)[order]
            self.qnext = np.concatenate(QNX)[order]
            self.isdel = np.concatenate(DEL)[order]
            self.isskip = np.concatenate(SKIP)[order]
            self.indel = np.concatenate(IND)[order]
        else:
            z = np.zeros(0, np.int64)
            self.p = self.ridx = self.qpn = self.qnext = self.indel = z
            self.ch = np.zeros(0, np.uint8)
            self.isdel = self.isskip = np.zeros(0, bool)

    def segment(self, i):
        s = self._segs.get(i)
        if s is None:
            s = self._segs[i] = AlignedSegment(self.rs, i)
        return s

    def columns(self):
        if len(self.p) == 0:
            return
        upos, first = np.unique(self.p, return_index=True)
        bounds = np.concatenate([first, [len(self.p)]])
        for k in range(len(upos)):
            yield PileupColumn(self, int(upos[k]), int(bounds[k]), int(bounds[k + 1]))


class Samfile:
    def __init__(self, path, mode=None, reference_filename=None, **kw):
        if path not in _REGISTRY:
            raise IOError("shim pysam: no ReadSet registered under %r" % path)
        self._c = _REGISTRY[path]

    @property
    def references(self):
        return tuple(self._c["order"])

    def is_valid_reference_name(self, name):
        return name in self._c["contigs"]

    def get_reference_length(self, name):
        return self._c["contigs"][name].contig_len

    def fetch(self, contig=None, start=None, end=None, multiple_iterators=False, **kw):
        rs = self._c["contigs"][contig]
        start = 0 if start is None else start
        end = rs.contig_len if end is None else end
        idx = np.nonzero((rs.pos < end) & (rs.ref_end > start) & ((rs.flag & 0x4) == 0))[0]
        for i in idx:
            yield AlignedSegment(rs, int(i))

    def pileup(self, contig=None, start=None, end=None, min_base_quality=13, flag_filter=0x704,
               truncate=False, multiple_iterators=False, **kw):
        rs = self._c["contigs"][contig]
        start = 0 if start is None else max(0, start)
        end = rs.contig_len if end is None else end
        if not truncate:
            raise NotImplementedError("shim pysam: only truncate=True is used by the reference")
        return _PileupEngine(rs, start, end, flag_filter).columns()

    def close(self):
        pass


AlignmentFile = Samfile


class FastaFile:
    def __init__(self, path, **kw):
        if path not in _REGISTRY:
            raise IOError("shim pysam: no reference registered under %r" % path)
        self._c = _REGISTRY[path]

    @property
    def references(self):
        return tuple(self._c["order"])

    def get_reference_length(self, name):
        return self._c["contigs"][name].contig_len

    def fetch(self, reference=None, start=None, end=None, **kw):
        rs = self._c["contigs"][reference]
        start = 0 if start is None else max(0, start)
        end = rs.contig_len if end is None else min(end, rs.contig_len)
        return rs.ref_string(start, end)


class _BedRow(tuple):
    pass


def asBed():
    return "bed"


class TabixFile:
    def __init__(self, path, **kw):
        if path not in _BEDS:
            raise IOError("shim pysam: no BED registered under %r" % path)
        self._b = _BEDS[path]

    def fetch(self, reference=None, start=None, end=None, parser=None):
        if reference not in self._b:
            raise ValueError("could not create iterator for region '%s'" % reference)
        return [_BedRow((reference, str(s), str(e))) for s, e in self._b[reference]]


class VariantFile:
    """Imported (never used) by generate_indel_pileups.py:4."""

    def __init__(self, *a, **kw):
        raise NotImplementedError("shim pysam: VariantFile is not used on the hot path")
