"""Golden-case definitions shared by make_golden.py (which runs the UNMODIFIED reference over
oracle/shim in the build container) and by the tests (which replay the same inputs through the
oracle and the CUDA path).  Inputs are regenerated from seeds; each fixture stores an input checksum
so generator drift is detected instead of silently comparing different data."""
import numpy as np

from nanocaller_b200.host.readset import ReadSet
from nanocaller_b200.synth import make_world

BASE_DCT = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1,
                seq="ont", supplementary=False, exclude_bed=None)


def _handmade():
    """Tiny hand-written alignments exercising CIGAR corner cases: soft/hard clips, padding, =/X,
    consecutive I and D, deletion spanning a candidate, N bases, reverse strand, junk flags,
    lower-case reference, a read ending exactly on a site."""
    rng = np.random.RandomState(5)
    ref = "".join("ACGT"[i] for i in rng.randint(0, 4, 400))
    ref = ref[:120] + ref[120:140].lower() + ref[140:]           # soft-masked run
    ref = ref[:300] + "N" + ref[301:]
    def mut(s, p, b):
        return s[:p] + b + s[p + 1:]
    recs = []
    # variant sites: 50 (het, linked), 90 (het), 130 (masked ref), 200 (hom), 260 (het), 300 (N ref)
    alt = {50: "T" if ref[50] != "T" else "A", 90: "G" if ref[90] != "G" else "C", 130: "A" if ref[130].upper() != "A" else "C",
           200: "C" if ref[200] != "C" else "G", 260: "A" if ref[260] != "A" else "T", 300: "A"}
    def read(start, end, hap, flag=0, lclip=0, rclip=0, extra=None):
        s = ref[start:end].upper().replace("N", "A")
        for p, b in alt.items():
            if start <= p < end and (hap == 1 or p == 200):
                s = mut(s, p - start, b)
        cig = "%dM" % (end - start)
        if extra == "eqx":
            cig = "%d=1X%d=" % (10, end - start - 11)
        elif extra == "del50" and start < 48 and end > 55:   # deletion covering site 50 (positions 49..52)
            k = 49 - start
            s = s[:k] + s[k + 4:]
            cig = "%dM4D%dM" % (k, end - start - k - 4)
        elif extra == "ins":                                 # 2I then 1I separated by P, then 2D 1D merged
            k = 20
            s = s[:k] + "GG" + "T" + s[k:k + 5] + s[k + 8:]
            cig = "%dM2I1P1I5M2D1D%dM" % (k, end - start - k - 8)
        elif extra == "nbase":
            s = mut(s, 90 - start, "N") if start <= 90 < end else s
        if lclip:
            s = "A" * lclip + s
            cig = "%dS" % lclip + cig
        if rclip:
            s = s + "C" * rclip
            cig = cig + "%dS" % rclip
        if extra == "hard":
            cig = "5H" + cig + "3H"
        return (start, flag, cig, s, hap + 1, 1)
    k = 0
    for start in range(0, 200, 9):
        end = min(400, start + 180 + (start * 7) % 40)
        hap = k & 1
        flag = 0x10 if (k % 3 == 0) else 0
        extra = {2: "eqx", 3: "del50", 5: "ins", 7: "nbase", 8: "hard"}.get(k % 11)
        if extra == "del50" and not (start < 48 and end > 55):
            extra = None
        recs.append(read(start, end, hap, flag, lclip=(k % 4 == 1) * 3, rclip=(k % 5 == 2) * 4, extra=extra))
        if k % 6 == 4:   # junk-flagged twins that must be ignored by the pileup
            recs.append(read(start, end, 1 - hap, flag | [0x100, 0x800, 0x400, 0x200, 0x4][k % 5]))
        k += 1
    recs.append(read(10, 91, 1))        # ends exactly on site 90 (last base is the site)
    recs.append(read(90, 300, 0))       # starts exactly on site 90
    recs.sort(key=lambda r: r[0])
    return ReadSet.from_records("tiny", ref, recs)


def _world(**kw):
    return make_world(**kw).reads


CASES = {
    # name: (readset factory, dct overrides, regions [(chrom,start,end,ploidy)], cpu for get_chunks, bed)
    "ont_diploid": (lambda: _world(chrom="chr20", preset="ont", contig_len=200_000, seed=20, coverage=30.0),
                    {}, [("chr20", 1, 200_000, "diploid")], 3, None),
    "ont_subregion_bed": (lambda: _world(chrom="chr20", preset="ont", contig_len=160_000, seed=21, coverage=24.0,
                                         mask_every=20000, mask_len=300, junk_frac=0.05, nbase_rate=0.002),
                          {"min_nbr_sites": 12, "exclude_bed": "mem://bed"}, [("chr20", 60_001, 100_000, "diploid")], 2,
                          {"chr20": [(61_000, 63_500), (90_000, 90_001)]}),
    "hifi_pacbio": (lambda: _world(chrom="chr1", preset="hifi", contig_len=120_000, seed=22, coverage=35.0),
                    {"seq": "pacbio", "threshold": [0.3, 0.7]}, [("chr1", 1, 120_000, "diploid")], 2, None),
    "short_ont": (lambda: _world(chrom="chr2", preset="short_ont", contig_len=100_000, seed=23, coverage=28.0),
                  {"seq": "short_ont", "threshold": [0.3, 0.7]}, [("chr2", 1, 100_000, "diploid")], 2, None),
    "ul_ont": (lambda: _world(chrom="chr3", preset="ont", contig_len=260_000, seed=24, coverage=20.0,
                              len_median=40000.0, len_max=150000, het_every=4000, sys_per_10k=30),
               {"seq": "ul_ont"}, [("chr3", 100_001, 160_000, "diploid")], 1, None),
    "ul_ont_extreme": (lambda: _world(chrom="chr3", preset="ont", contig_len=260_000, seed=24, coverage=20.0,
                                      len_median=40000.0, len_max=150000, het_every=4000, sys_per_10k=30),
                       {"seq": "ul_ont_extreme"}, [("chr3", 100_001, 160_000, "diploid")], 1, None),
    "haploid": (lambda: _world(chrom="chrY", preset="ont", contig_len=120_000, seed=25, coverage=30.0, ploidy=1),
                {}, [("chrY", 1, 120_000, "haploid")], 2, None),
    "lowcov": (lambda: _world(chrom="chr4", preset="ont", contig_len=80_000, seed=26, coverage=5.0),
               {"mincov": 6}, [("chr4", 1, 80_000, "diploid")], 1, None),
    "handmade": (_handmade, {"mincov": 2, "threshold": [0.3, 0.7]}, [("tiny", 1, 400, "diploid")], 1, None),
    "empty": (lambda: _world(chrom="chr5", preset="ont", contig_len=30_000, seed=27, coverage=30.0, het_every=0,
                             hom_every=0, sys_per_10k=0, sub_rate=0.0),
              {}, [("chr5", 1, 30_000, "diploid")], 1, None),
}


def case_inputs(name):
    factory, over, regions, cpu, bed = CASES[name]
    dct = dict(BASE_DCT)
    dct.update(over)
    return factory(), dct, regions, cpu, bed
