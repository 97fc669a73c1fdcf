#!/usr/bin/env python
"""Indel build on a synthetic HiFi contig (261-column windows): device time of the star alignment with the paired 16-bit kernel
(nine-column strips) and, NC_INDEL_ALIGN_SCALAR=1, with the one-slice-per-warp kernel.  python tools/hifi_align_probe.py [Mb]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from nanocaller_b200.host import indel_pileups, snp_pileups
    from nanocaller_b200.synth import make_world
    mb = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    L = int(mb * 1e6)
    rs = make_world(chrom="chrH", preset="hifi", contig_len=L, seed=5, coverage=35.0, indel_every=2000, indel_maxlen=50).reads
    dct = dict(mincov=4, maxcov=160, seq="pacbio", del_t=0.4, ins_t=0.4, impute_indel_phase=False, supplementary=False, win_size=40, small_win_size=4)
    chunks = [{"chrom": "chrH", "start": s, "end": min(L, s + 100_000 - 1), "ploidy": "diploid"} for s in range(1, L, 100_000)]
    ctx = snp_pileups.context(0)
    out = {}
    for mode in ("0", "1"):
        os.environ["NC_INDEL_ALIGN_SCALAR"] = mode
        for _ in range(3):
            snp_pileups._staged.clear()
            meta, _, _ = indel_pileups.scan_build(ctx, rs, dct, chunks, want_tensors=False)
        tm = ctx.indel_timings()
        out["scalar" if mode == "1" else "paired"] = {"align_ms": tm["align_ms"], "sites": int(len(meta)), "slices": int(tm.get("n_entries", 0))}
    out["contig_mb"] = mb
    print(json.dumps(out))


if __name__ == "__main__":
    main()
