"""Generate SNP record fixtures by running the UNMODIFIED reference worker `nanocaller_src.snpCaller.caller`
(snpCaller.py:57-198: model construction from model_architect*.py, batching, coverage scaling, genotype decision, VCF record
text) over oracle/shim — pysam, intervaltree and the stand-in tensorflow, whose primitive ops (conv / dense / selu / softmax) are
float32 torch while everything else is the reference's own code — with the released weights in /root/reference.
The worker is started as a child process exactly as `call_manager` starts it (snpCaller.py:238), because it names its output
file after `current_process()._identity`.

    python tests/golden/make_golden_records.py [case ...]      (build container only)
Writes tests/golden/records_<case>.vcf.txt (the worker's intermediate VCF, record lines in the order written).
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

import pysam  # the shim  # noqa: E402
from nanocaller_src import snpCaller  # noqa: E402  (reference, unchanged)
from nanocaller_src.utils import get_chunks  # noqa: E402  (reference, unchanged)
from tests.golden.cases import case_inputs  # noqa: E402

RECORD_CASES = {"ont_diploid": {}, "haploid": {}, "ont_subregion_bed": {"disable_coverage_normalization": True}, "lowcov": {},
                "hifi_pacbio": {"snp_model": "CCS-HG002"},
                # haploid contig with --disable_coverage_normalization (snpCaller.py:169-170: hap_train_coverage / dp per site)
                "haploid_nonorm": {"_case": "haploid", "disable_coverage_normalization": True}}


def run_case(name, over):
    over = dict(over)
    rs, dct, regions, cpu, bed = case_inputs(over.pop("_case", name))
    pysam.unregister_all()
    pysam.register("mem://bam", rs)
    if bed is not None:
        pysam.register_bed("mem://bed", bed)
    tmp = tempfile.mkdtemp(prefix="nc_rec_")
    params = dict(dct, sam_path="mem://bam", fasta_path="mem://bam", snp_model="ONT-HG002", prefix="g",
                  intermediate_snp_files_dir=tmp, disable_coverage_normalization=False)
    params.update(over)
    chunks = get_chunks(regions, cpu)
    ctx = mp.get_context("fork")
    mgr = ctx.Manager()
    chunks_q, counter_q, files = mgr.Queue(), mgr.Queue(), mgr.list()
    for ch in chunks:
        chunks_q.put(ch)
    p = ctx.Process(target=snpCaller.caller, args=(params, chunks_q, counter_q, files))
    p.start()
    p.join()
    assert p.exitcode == 0, p.exitcode
    assert len(files) == 1
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
