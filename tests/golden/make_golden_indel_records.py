"""Generate indel record fixtures by running the UNMODIFIED reference worker `nanocaller_src.indelCaller.indel_run`
(indelCaller.py:41-189: model construction from model_architect_indel*.py, hstack of the three tensors, batching, genotype decision,
`prev` overlap suppression, VCF record text) over oracle/shim — pysam, parasail and muscle stand-ins as in make_golden_indel.py plus
the stand-in tensorflow (primitive ops only) — with the released weights in /root/reference.  The worker runs as a child process
like `call_manager` starts it (indelCaller.py:346).

    python tests/golden/make_golden_indel_records.py [case ...]      (build container only)
Writes tests/golden/records_<case>.vcf.txt.
"""
import multiprocessing as mp
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
os.environ["PATH"] = os.path.join(ROOT, "oracle", "shim", "bin") + os.pathsep + os.environ["PATH"]

import pysam  # the shim  # noqa: E402
from nanocaller_src import indelCaller  # noqa: E402  (reference, unchanged)
from nanocaller_src.utils import get_chunks  # noqa: E402
from tests.golden.indel_cases import indel_case_inputs  # noqa: E402

RECORD_CASES = {"indel_ont": "ONT-HG002", "indel_haploid": "ONT-HG002", "indel_impute_hifi": "CCS-HG002", "indel_sub": "ONT-HG002",
                "indel_impute_ont": "ONT-HG002"}


def run_case(name, model):
    rs, dct, regions, cpu, mcs = indel_case_inputs(name)
    pysam.unregister_all()
    pysam.register("mem://bam", rs)
    tmp = tempfile.mkdtemp(prefix="nc_irec_")
    params = dict(dct, fasta_path="mem://bam", indel_model=model, prefix="g", intermediate_indel_files_dir=tmp)
    chunks = get_chunks(regions, cpu, max_chunk_size=mcs)
    ctx = mp.get_context("fork")
    mgr = ctx.Manager()
    job_q, counter_q, files, indel_dict = mgr.Queue(), mgr.Queue(), mgr.list(), mgr.dict()
    for ch in chunks:
        job_q.put(("indel", dict(ch, sam_path="mem://bam")))
    p = ctx.Process(target=indelCaller.indel_run, args=(params, indel_dict, job_q, counter_q, files))
    p.start()
    p.join()
    assert p.exitcode == 0, p.exitcode
    text = open(files[0]).read()
    shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(HERE, "records_%s.vcf.txt" % name), "w") as f:
        f.write("# %s\n" % rs.checksum())
        f.write(text)
    print("  %s: %d chunks, %d records" % (name, len(chunks), text.count("\n")), flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(RECORD_CASES)):
        print("case", nm, flush=True)
        run_case(nm, RECORD_CASES[nm])
