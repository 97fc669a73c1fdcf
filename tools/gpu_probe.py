"""Quick GPU timing probe (development aid): synthetic world -> stage -> scan -> forward, prints timings."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from nanocaller_b200.host import capi, snp_pileups, weights as W   # noqa: E402
from nanocaller_b200.synth import make_world                       # noqa: E402
from oracle.snp_oracle import get_chunks                           # noqa: E402

L = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t = time.time()
rs = make_world(chrom="chr20", preset="ont", contig_len=L, seed=20, coverage=30.0).reads
print("world %.1fs reads=%d aligned=%.3g cigar=%d seq4=%d" % (time.time() - t, rs.n, rs.aligned_bases(), len(rs.cigar), len(rs.seq4)), flush=True)
dct = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont", supplementary=False)
chunks = get_chunks([("chr20", 1, L, "diploid")], 1)
ctx = capi.Context(0)
tensors, meta = W.load_model("snp", "ONT-HG002")
ctx.load_snp_weights(W.pack_snp_blob(tensors, False), meta["train_coverage"], False)
for it in range(3):
    t0 = time.time()
    ctx.stage_reads(rs)
    ctx.sync()
    t1 = time.time()
    n = snp_pileups.scan_chunks(ctx, rs, dct, chunks, "diploid") if it == 0 else ctx.snp_scan(capi.snp_params(dct, "diploid"), [(c["start"], c["end"]) for c in chunks])
    ctx.sync()
    t2 = time.time()
    probs = ctx.snp_forward(True, impl)
    t3 = time.time()
    tm = ctx.timings()
    print("iter %d sites=%d stage %.1fms scan(wall) %.1fms fwd(wall) %.1fms | dev: decode %.3f scan %.3f tensor %.3f cnn %.3f ms launches %d" % (
        it, n, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, tm["decode_ms"], tm["scan_ms"], tm["tensor_ms"], tm["cnn_ms"], tm["launches"]), flush=True)
print("probs mean", probs.mean(0), "sites/s(dev)", n / ((tm["decode_ms"] + tm["scan_ms"] + tm["tensor_ms"] + tm["cnn_ms"]) * 1e-3))
