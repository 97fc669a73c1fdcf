"""ORACLE / TEST INFRASTRUCTURE ONLY — minimal stand-in for the `intervaltree` package so the
reference's `generate_SNP_pileups.py` imports unchanged.  Implements what the reference uses
(`generate_SNP_pileups.py:116-119`): construction from `Interval(begin, end, data)` and the
point query `tree.overlaps(pos)`  <=>  any begin <= pos < end.  Null intervals raise ValueError
like the real package."""
from collections import namedtuple
import bisect

Interval = namedtuple("Interval", ["begin", "end", "data"], defaults=[None])


class IntervalTree:
    def __init__(self, intervals=None):
        ivs = sorted(set(intervals)) if intervals is not None else []
        for iv in ivs:
            if iv.begin >= iv.end:
                raise ValueError("IntervalTree: Null Interval objects not allowed in IntervalTree: %r" % (iv,))
        self._begins = [iv.begin for iv in ivs]
        # running maximum of ends lets a point query be answered with one bisect
        self._maxend = []
        m = None
        for iv in ivs:
            m = iv.end if m is None or iv.end > m else m
            self._maxend.append(m)

    def overlaps(self, begin, end=None):
        if end is not None:
            raise NotImplementedError
        k = bisect.bisect_right(self._begins, begin)
        return k > 0 and self._maxend[k - 1] > begin

    def __len__(self):
        return len(self._begins)
