"""Generate the golden fixtures by running the UNMODIFIED reference modules
(/root/reference/nanocaller_src/generate_SNP_pileups.py, utils.py) over oracle/shim.

Runs only in the build container (needs /root/reference).  Usage:
    python tests/golden/make_golden.py [case ...]
Writes tests/golden/<case>.npz.
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

import pysam  # the shim  # noqa: E402
from nanocaller_src.generate_SNP_pileups import get_snp_testing_candidates  # noqa: E402  (reference, unchanged)
from nanocaller_src.utils import get_chunks  # noqa: E402  (reference, unchanged)
from tests.golden.cases import CASES, case_inputs  # noqa: E402


def run_case(name):
    rs, dct, regions, cpu, bed = case_inputs(name)
    pysam.unregister_all()
    pysam.register("mem://bam", rs)
    if bed is not None:
        pysam.register_bed("mem://bed", bed)
    d = dict(dct, sam_path="mem://bam", fasta_path="mem://bam")
    chunks = get_chunks(regions, cpu)
    out = {"input_checksum": np.array(rs.checksum()), "n_chunks": np.array(len(chunks)),
           "chunks_json": np.array(json.dumps(chunks))}
    for ci, chunk in enumerate(chunks):
        t = time.time()
        pos, ref, mat, dp, freq, depth, fwd, rev = get_snp_testing_candidates(d, chunk)
        n = len(pos)
        print("  %s chunk %d %s: %d candidates, depth %.4f (%.1fs)" % (name, ci, chunk, n, depth, time.time() - t), flush=True)
        out["c%d_pos" % ci] = np.asarray(pos, np.int64)
        out["c%d_ref" % ci] = np.asarray(ref, np.int8).reshape(n, 4) if n else np.zeros((0, 4), np.int8)
        m = np.asarray(mat, np.float32).reshape(n, 5, 41, 5) if n else np.zeros((0, 5, 41, 5), np.float32)
        assert np.all(m == np.round(m)) and np.abs(m).max(initial=0) < 32768
        out["c%d_mat" % ci] = m.astype(np.int16)
        out["c%d_dp" % ci] = np.asarray(dp, np.int64)
        out["c%d_freq" % ci] = np.asarray(freq, np.float64)
        out["c%d_depth" % ci] = np.asarray(depth, np.float64)
        out["c%d_fwd" % ci] = np.asarray(fwd, np.float64).reshape(n, 4).astype(np.int16) if n else np.zeros((0, 4), np.int16)
        out["c%d_rev" % ci] = np.asarray(rev, np.float64).reshape(n, 4).astype(np.int16) if n else np.zeros((0, 4), np.int16)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        print("case", nm, flush=True)
        run_case(nm)
