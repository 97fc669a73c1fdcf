#!/usr/bin/env python
"""One synthetic genome through the command line (`nanocaller_b200.cli`), on 1 GPU or sharded over the GPUs of the box under
torchrun (host/multi.py: contigs or chunk runs per rank, NCCL gather of the ranks' record text to rank 0), and a byte-for-byte
comparison of the two runs' VCF files.  BASELINE.json configs[3] (mode=all, HiFi `ccs` preset, 35x) and configs[4]
(`--haploid_genome`, ONT preset, 60x), scaled: 24 contigs with GRCh38's length proportions summing to --mb megabases.

    python tools/multi_gpu_genome.py --config hifi --mb 300 --out /tmp/g1                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_genome.py --config hifi --mb 300 --out /tmp/g8 --compare /tmp/g1            # 8 GPUs + comparison

The reads are generated in memory (seeded, csrc/synth.cpp), every rank only the contigs it owns: the path under test is
region sharding + kernels + gather + VCF writing, not BGZF inflate (bench.py --from-bam measures that).  Prints one JSON line.
"""
import argparse
import gzip
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GRCH38_MB = [("chr1", 248), ("chr2", 242), ("chr3", 198), ("chr4", 190), ("chr5", 181), ("chr6", 171), ("chr7", 159), ("chr8", 145),
             ("chr9", 138), ("chr10", 133), ("chr11", 135), ("chr12", 133), ("chr13", 114), ("chr14", 107), ("chr15", 102), ("chr16", 90),
             ("chr17", 83), ("chr18", 80), ("chr19", 58), ("chr20", 64), ("chr21", 47), ("chr22", 51), ("chrX", 156), ("chrY", 57)]
CONFIGS = {
    # configs[3]: mode=all, HiFi preset (NanoCaller:74: pacbio, CCS-HG002 models, 0.3,0.7, ins/del 0.4, impute_indel_phase), 35x
    "hifi": dict(argv=["--preset", "ccs", "--mode", "all"], synth=dict(preset="hifi", coverage=35.0, indel_every=2000, indel_maxlen=50)),
    # configs[4]: --haploid_genome, ONT preset, 60x (hom-alt truth only)
    "haploid": dict(argv=["--preset", "ont", "--haploid_genome", "--mode", "all"],
                    synth=dict(preset="ont", coverage=60.0, ploidy=1, indel_every=2000, indel_maxlen=50)),
    # configs[1]-shaped: one long contig, mode=snps: chunk runs (not contigs) are sharded
    "snps": dict(argv=["--preset", "ont", "--mode", "snps"], synth=dict(preset="ont", coverage=30.0)),
}


GEN = {"seconds": 0.0, "bases": 0}


def genome(config, mb):
    from nanocaller_b200.synth import make_world
    total = sum(n for _, n in GRCH38_MB)
    contigs = {}
    table = GRCH38_MB if config != "snps" else [("chr20", total)]
    for i, (name, n) in enumerate(table):
        length = max(200_000, int(round(n / total * mb * 1e6)))
        kw = dict(CONFIGS[config]["synth"])
        preset = kw.pop("preset")

        def factory(name=name, length=length, i=i, kw=kw, preset=preset):
            t = time.time()
            rs = make_world(chrom=name, preset=preset, contig_len=length, seed=1000 + i, **kw).reads
            GEN["seconds"] += time.time() - t
            GEN["bases"] += rs.aligned_bases()
            return rs
        contigs[name] = (length, factory)
    return contigs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=sorted(CONFIGS), default="hifi")
    ap.add_argument("--mb", type=float, default=300.0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--compare", default=None, help="output directory of another run of the same genome: compare the VCF files byte for byte")
    ap.add_argument("--cpu", type=int, default=16)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    from nanocaller_b200 import cli
    from nanocaller_b200.host import sources
    sources.register_lazy("mem://genome", genome(a.config, a.mb))
    argv = ["--bam", "mem://genome", "--ref", "mem://genome", "--output", a.out, "--cpu", str(a.cpu), "--suppress_progress_bar"] + CONFIGS[a.config]["argv"]
    t0 = time.time()
    out = cli.main(argv)
    dt = time.time() - t0
    gen = [GEN["seconds"], float(GEN["bases"])]
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor(gen, dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
        ts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        gen = [max(x[0].item() for x in ts), sum(x[1].item() for x in ts)]
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    res = {"config": a.config, "megabases": a.mb, "n_gpus": world, "seconds": dt, "read_generation_seconds_max_rank": gen[0],
           "seconds_without_read_generation": dt - gen[0], "aligned_bases": gen[1], "contigs": len(genome(a.config, a.mb)),
           "sharding": out.get("sharding", "single process"), "records": {k: v for k, v in out.items() if k.startswith("n_")},
           "seconds_by_stage": {k: v for k, v in out.items() if k.endswith("_seconds")}}
    if a.compare:
        same = {}
        for name in sorted(os.listdir(a.out)):
            if name.endswith(".vcf.gz") and os.path.exists(os.path.join(a.compare, name)):
                x = gzip.open(os.path.join(a.out, name), "rb").read()
                y = gzip.open(os.path.join(a.compare, name), "rb").read()
                same[name] = {"identical": x == y, "bytes": len(x), "records": x.count(b"\n") - sum(1 for ln in x.split(b"\n") if ln.startswith(b"#"))}
        res["compared_with"] = a.compare
        res["vcf_identical"] = same
        res["all_identical"] = bool(same) and all(v["identical"] for v in same.values())
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
