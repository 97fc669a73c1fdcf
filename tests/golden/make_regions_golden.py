"""Run the UNMODIFIED reference `nanocaller_src.utils.get_regions_list` (utils.py:6-65) over the pysam shim on a set of argument
scenarios and store its answers in tests/golden/reference_regions.json.  Build container only."""
import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")

import pysam  # the shim  # noqa: E402
from nanocaller_src.utils import get_regions_list  # noqa: E402  (reference, unchanged)
from nanocaller_b200.host.readset import ReadSet  # noqa: E402

CONTIGS = {"named": [("chr1", 5000), ("chr2", 4000), ("chr3", 3000), ("chrX", 2500), ("chrY", 800), ("chrM", 160), ("chrUn_1", 90)],
           "plain": [("1", 5000), ("2", 4000), ("X", 2500), ("Y", 800), ("M", 160)]}
BED = "chr2\t10\t900\nchrQ\t1\t5\nchr1\t100\t200\tname\nchrY\t3\t30\n"
SCENARIOS = [
    dict(name="all_contigs", source="named"),
    dict(name="all_contigs_haploid_genome", source="named", haploid_genome=True),
    dict(name="all_contigs_haploid_X", source="named", haploid_X=True),
    dict(name="wgs_chr", source="named", wgs_contigs="chr1-22XY"),
    dict(name="wgs_plain", source="plain", wgs_contigs="1-22XY", haploid_X=True),
    dict(name="wgs_mismatch", source="plain", wgs_contigs="chr1-22XY"),
    dict(name="regions_mixed", source="named", regions=["chr1", "chr2:100-5000", "chrQ", "chr3:5", "chrX", "chr1:7-9:3", "chrM"], haploid_X=True),
    dict(name="regions_plain", source="plain", regions=["X:5-50", "2"], haploid_genome=True),
    dict(name="regions_none_valid", source="named", regions=["chrQ", "chr9:4"]),
    dict(name="bed", source="named", bed=True),
    dict(name="wgs_wins_over_regions", source="named", wgs_contigs="chr1-22XY", regions=["chr2:1-5"]),
]


def _readset(name, length):
    z = np.zeros(0)
    return ReadSet(name, np.full(length, ord("A"), np.uint8), z, z, np.zeros(1), z, np.zeros(1), z, z)


def main():
    out = {"contigs": CONTIGS, "bed": BED, "scenarios": []}
    bed_path = os.path.join(HERE, "_regions.bed")
    open(bed_path, "w").write(BED)
    for sc in SCENARIOS:
        pysam.unregister_all()
        pysam.register("mem://r", [_readset(n, l) for n, l in CONTIGS[sc["source"]]])
        args = argparse.Namespace(bam="mem://r", wgs_contigs=sc.get("wgs_contigs"), regions=sc.get("regions"), bed=bed_path if sc.get("bed") else None,
                                  haploid_genome=sc.get("haploid_genome", False), haploid_X=sc.get("haploid_X", False))
        buf = io.StringIO()
        try:
            with contextlib.redirect_stdout(buf):
                res = [list(r) for r in get_regions_list(args)]
        except SystemExit as e:
            res = {"exit": e.code}
        msgs = [ln.split(": ", 1)[1] for ln in buf.getvalue().splitlines() if ": " in ln]
        out["scenarios"].append(dict(sc, result=res, messages=msgs))
        print(sc["name"], res if isinstance(res, dict) else len(res), msgs)
    os.remove(bed_path)
    json.dump(out, open(os.path.join(HERE, "reference_regions.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
