"""In-memory alignment store for one contig, in BAM-native encoding.

This is the host-side staging format: what the BAM reader produces, what the synthetic generator
produces, what the device stager uploads (`nc_stage_reads`) and what the CPU oracle piles up.
Arrays mirror the BAM record fields (SAM spec §4.2): 0-based `pos`, `flag`, CIGAR as
`len<<4|op` u32 words with op codes MIDNSHP=X = 0..8, bases 4 bits each ("=ACMGRSVTWYHKDBN"),
every read's packed sequence starting on a byte boundary.
"""
import re

import numpy as np

CIGAR_OPS = "MIDNSHP=X"
_REF_CONSUME = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], dtype=np.int64)
_QRY_CONSUME = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], dtype=np.int64)
NIB_CHARS = "=ACMGRSVTWYHKDBN"
_NIB_OF = {c: i for i, c in enumerate(NIB_CHARS)}
_NIB_LUT = np.frombuffer(NIB_CHARS.encode(), dtype=np.uint8)


class ReadSet:
    def __init__(self, chrom, ref, pos, flag, cigar_off, cigar, seq_off, l_seq, seq4,
                 hp=None, ps=None, qnames=None):
        self.chrom = chrom
        self.ref = np.ascontiguousarray(ref, dtype=np.uint8)
        self.pos = np.ascontiguousarray(pos, dtype=np.int32)
        self.flag = np.ascontiguousarray(flag, dtype=np.uint16)
        self.cigar_off = np.ascontiguousarray(cigar_off, dtype=np.int64)
        self.cigar = np.ascontiguousarray(cigar, dtype=np.uint32)
        self.seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
        self.l_seq = np.ascontiguousarray(l_seq, dtype=np.int32)
        self.seq4 = np.ascontiguousarray(seq4, dtype=np.uint8)
        n = len(self.pos)
        self.hp = np.zeros(n, np.int8) if hp is None else np.ascontiguousarray(hp, dtype=np.int8)
        self.ps = np.zeros(n, np.int32) if ps is None else np.ascontiguousarray(ps, dtype=np.int32)
        self.qnames = qnames
        assert len(self.cigar_off) == n + 1 and len(self.seq_off) == n + 1
        if n and np.any(np.diff(self.pos) < 0):
            raise ValueError("reads must be coordinate-sorted")
        self._ref_end = None

    # ---- derived ----
    @property
    def n(self):
        return len(self.pos)

    @property
    def contig_len(self):
        return len(self.ref)

    @property
    def ref_end(self):
        """Exclusive 0-based reference end of every read (pos + reference-consuming CIGAR length)."""
        if self._ref_end is None:
            ops = self.cigar & 15
            lens = (self.cigar >> 4).astype(np.int64)
            rl = lens * _REF_CONSUME[ops]
            cs = np.concatenate([[0], np.cumsum(rl)])
            span = cs[self.cigar_off[1:]] - cs[self.cigar_off[:-1]]
            self._ref_end = (self.pos.astype(np.int64) + span).astype(np.int32)
        return self._ref_end

    def qname(self, i):
        return self.qnames[i] if self.qnames is not None else "r%08d" % i

    def read_cigar(self, i):
        return self.cigar[self.cigar_off[i]:self.cigar_off[i + 1]]

    def read_nibbles(self, i):
        """Unpacked 4-bit base codes of read i (length l_seq)."""
        b = self.seq4[self.seq_off[i]:self.seq_off[i + 1]]
        nib = np.empty(len(b) * 2, np.uint8)
        nib[0::2] = b >> 4
        nib[1::2] = b & 15
        return nib[:self.l_seq[i]]

    def query_sequence(self, i):
        return _NIB_LUT[self.read_nibbles(i)].tobytes().decode()

    def ref_string(self, start, end):
        """Reference characters of 0-based [start, end)."""
        return self.ref[max(0, start):max(0, end)].tobytes().decode()

    def aligned_bases(self):
        ops = self.cigar & 15
        return int(((self.cigar >> 4).astype(np.int64) * (_REF_CONSUME[ops] & _QRY_CONSUME[ops])).sum())

    def checksum(self):
        """Order-sensitive 64-bit checksum of every input array (pins generator determinism in fixtures)."""
        import hashlib
        h = hashlib.sha256()
        for a in (self.ref, self.pos, self.flag, self.cigar_off, self.cigar, self.seq_off, self.l_seq,
                  self.seq4, self.hp, self.ps):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()[:16]

    # ---- construction from text records (tests) ----
    @staticmethod
    def from_records(chrom, ref, records):
        """records: iterable of dicts/tuples (pos0, flag, cigar_string, seq_string[, hp, ps, qname]); sorted by pos."""
        pos, flag, l_seq, hp, ps, qn = [], [], [], [], [], []
        cig_words, cig_off, seq_bytes, seq_off = [], [0], [], [0]
        for rec in records:
            p, f, cg, sq = rec[0], rec[1], rec[2], rec[3]
            h = rec[4] if len(rec) > 4 else 0
            s = rec[5] if len(rec) > 5 else 0
            q = rec[6] if len(rec) > 6 else None
            words = [(int(l) << 4) | CIGAR_OPS.index(o) for l, o in re.findall(r"(\d+)([MIDNSHP=X])", cg)]
            qlen = sum((w >> 4) for w in words if _QRY_CONSUME[w & 15])
            if qlen != len(sq):
                raise ValueError("CIGAR query length %d != sequence length %d" % (qlen, len(sq)))
            nib = [_NIB_OF.get(c.upper(), 15) for c in sq]
            if len(nib) & 1:
                nib.append(0)
            packed = [(nib[k] << 4) | nib[k + 1] for k in range(0, len(nib), 2)]
            pos.append(p); flag.append(f); l_seq.append(len(sq)); hp.append(h); ps.append(s); qn.append(q)
            cig_words += words; cig_off.append(len(cig_words))
            seq_bytes += packed; seq_off.append(len(seq_bytes))
        qnames = None if all(q is None for q in qn) else [q if q is not None else "r%08d" % i for i, q in enumerate(qn)]
        ref_arr = np.frombuffer(ref.encode(), dtype=np.uint8) if isinstance(ref, str) else ref
        return ReadSet(chrom, ref_arr, pos, flag, cig_off, np.array(cig_words, np.uint32), seq_off,
                       l_seq, np.array(seq_bytes, np.uint8), hp, ps, qnames)

    def window(self, lo0, hi0):
        """Reads that can overlap the 0-based half-open reference window [lo0, hi0), as a ReadSet over VIEWS of this one's arrays:
        the contiguous BAM-order range from the first read ending after lo0 to the last read starting before hi0 (reads in between
        that end before lo0 stay in — they overlap nothing in the window, and the range stays contiguous).  The reference sequence
        is kept whole, so coordinates do not change.  Used to hand one rank of a chunk-sharded run only its part of a contig."""
        after = np.nonzero(self.ref_end > lo0)[0]
        i0 = int(after[0]) if len(after) else self.n
        i1 = max(i0, int(np.searchsorted(self.pos, hi0, side="left")))
        c0, s0 = int(self.cigar_off[i0]), int(self.seq_off[i0])
        w = ReadSet(self.chrom, self.ref, self.pos[i0:i1], self.flag[i0:i1], self.cigar_off[i0:i1 + 1] - c0,
                    self.cigar[c0:int(self.cigar_off[i1])], self.seq_off[i0:i1 + 1] - s0, self.l_seq[i0:i1],
                    self.seq4[s0:int(self.seq_off[i1])], self.hp[i0:i1], self.ps[i0:i1],
                    None if self.qnames is None else self.qnames[i0:i1])
        w._ref_end = np.ascontiguousarray(self.ref_end[i0:i1])
        return w

    def subset(self, keep):
        """New ReadSet with reads where boolean mask `keep` is set (arrays re-packed)."""
        idx = np.nonzero(keep)[0]
        cl = (self.cigar_off[1:] - self.cigar_off[:-1])[idx]
        sl = (self.seq_off[1:] - self.seq_off[:-1])[idx]
        co = np.concatenate([[0], np.cumsum(cl)])
        so = np.concatenate([[0], np.cumsum(sl)])
        cig = np.concatenate([self.cigar[self.cigar_off[i]:self.cigar_off[i + 1]] for i in idx]) if len(idx) else np.zeros(0, np.uint32)
        sq = np.concatenate([self.seq4[self.seq_off[i]:self.seq_off[i + 1]] for i in idx]) if len(idx) else np.zeros(0, np.uint8)
        qn = [self.qnames[i] for i in idx] if self.qnames is not None else ["r%08d" % i for i in idx]
        return ReadSet(self.chrom, self.ref, self.pos[idx], self.flag[idx], co, cig, so, self.l_seq[idx], sq,
                       self.hp[idx], self.ps[idx], qn)
