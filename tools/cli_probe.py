"""Development aid: `python tools/cli_probe.py [contig_bp] [mode]` — synthetic BAM + FASTA on disk, then the command line end to end."""
import os
import sys
import tempfile
import time

sys.path.insert(0, ".")
from nanocaller_b200 import cli                      # noqa: E402
from nanocaller_b200.host import bamio               # noqa: E402
from nanocaller_b200.synth import make_world         # noqa: E402

L = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "all"
rs = make_world(chrom="chr20", preset="ont", contig_len=L, seed=20, coverage=30.0, indel_every=2000, indel_maxlen=50).reads
d = tempfile.mkdtemp(prefix="nc_cli_")
bam, fa = os.path.join(d, "x.bam"), os.path.join(d, "x.fa")
t = time.time()
bamio.write_bam(bam, [rs]); bamio.write_fasta(fa, [rs])
print("wrote %s (%.1f MB) in %.1fs" % (bam, os.path.getsize(bam) / 1e6, time.time() - t), flush=True)
for rep in range(2):
    t = time.time()
    out = cli.main(["--bam", bam, "--ref", fa, "--mode", mode, "--preset", "ont", "--output", os.path.join(d, "o%d" % rep), "--suppress_progress_bar"])
    print("run %d: total %.2fs read %.2fs snps %.2fs (%d records) indels %.2fs (%d records) launches %d" % (
        rep, time.time() - t, out.get("read_seconds", 0), out.get("snp_seconds", 0), out.get("n_snp_records", 0),
        out.get("indel_seconds", 0), out.get("n_indel_records", 0), out.get("launches", 0)), flush=True)
