"""Generate indel golden fixtures by running the UNMODIFIED reference module
/root/reference/nanocaller_src/generate_indel_pileups.py over oracle/shim (pysam, intervaltree, parasail stand-ins)
with oracle/shim/bin/muscle first on PATH.  Candidate positions, read slices, haplotype split, tensor assembly,
consensus and the allele logic are the reference's own code; the MSA and the pairwise alignment underneath are this
repo's stated stand-ins (oracle/star_msa.py) because MUSCLE / parasail cannot be had here.

    python tests/golden/make_golden_indel.py [case ...]      (build container only)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
os.environ["PATH"] = os.path.join(ROOT, "oracle", "shim", "bin") + os.pathsep + os.environ["PATH"]

import pysam  # the shim  # noqa: E402
from nanocaller_src.generate_indel_pileups import get_indel_testing_candidates  # noqa: E402  (reference, unchanged)
from nanocaller_src.generate_indel_pileups_haploid import get_indel_testing_candidates_haploid  # noqa: E402  (reference, unchanged)
from nanocaller_src.utils import get_chunks  # noqa: E402
from tests.golden.indel_cases import INDEL_CASES, indel_case_inputs  # noqa: E402


def run_case(name):
    rs, dct, regions, cpu, mcs = indel_case_inputs(name)
    pysam.unregister_all()
    pysam.register("mem://bam", rs)
    d = dict(dct, fasta_path="mem://bam")
    chunks = get_chunks(regions, cpu, max_chunk_size=mcs)
    out = {"input_checksum": np.array(rs.checksum()), "chunks_json": np.array(json.dumps(chunks))}
    for ci, chunk in enumerate(chunks):
        t = time.time()
        ch = dict(chunk, sam_path="mem://bam")
        if chunk["ploidy"] == "haploid":
            pos, x2, alleles = get_indel_testing_candidates_haploid(d, ch)
            x0 = x1 = x2
            phase = []
        else:
            pos, x0, x1, x2, alleles, phase = get_indel_testing_candidates(d, ch)
        n = len(pos)
        print("  %s chunk %d %s: %d candidates (%.1fs)" % (name, ci, chunk, n, time.time() - t), flush=True)
        out["c%d_pos" % ci] = np.asarray(pos, np.int64)
        for k, x in (("x0", x0), ("x1", x1), ("x2", x2)):
            a = np.asarray(x, np.float64).reshape(n, 5, 128, 2) if n else np.zeros((0, 5, 128, 2))
            assert np.array_equal(a, a.astype(np.float32).astype(np.float64))      # float32 values in a float64 container
            out["c%d_%s" % (ci, k)] = a.astype(np.float32)
        out["c%d_alleles" % ci] = np.array(json.dumps(alleles))
        out["c%d_phase" % ci] = np.array(json.dumps(phase))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(INDEL_CASES)):
        print("case", nm, flush=True)
        run_case(nm)
