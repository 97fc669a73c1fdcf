"""`python -m nanocaller_b200 --bam X.bam --ref X.fa --mode snps --preset ont ...`

The reference's command line (`NanoCaller:84-158`: same flags, defaults, preset table `:66-77`, region rules
`nanocaller_src/utils.py:6-65`, chunk grid `utils.py:67-83`, output names `snpCaller.py:251-252`,
`indelCaller.py:385-395`) driving the B200 path: BAM/FASTA -> staging arrays -> CUDA pileup + CNN -> VCF records ->
sorted BGZF VCFs.  Between the two stages the reference shells out to `whatshap phase/haplotag`
(`indelCaller.py:234-251`); here `host/phasing.py` (libnc_phase.so, an own algorithm) phases the PASS SNP calls and tags
the reads in memory — in `--mode all` when the contig's reads carry no HP tags yet, or whenever `--phase` is given;
reads that are already haplotagged are used as they are (as `--mode indels` does in the reference).  Not here:
`--enable_whatshap`'s genotype revision.  `rtg vcfdecompose | vcffilter --non-snps-only` (`indelCaller.py:391`) has an own-rule
stand-in behind `--decompose_indels` (`host/vcf_decompose.py`, off by default).  `--cpu` only sets
the chunk grid (and with it the normalisation groups); the work runs on the GPU of
`--device`.  There is no CPU fallback: without an sm_100 GPU the run fails in `nc_create`."""
import argparse
import datetime
import gzip
import os
import sys
import time

PRESETS = {      # NanoCaller:66-77
    "ont": dict(sequencing="ont", snp_model="ONT-HG002", indel_model="ONT-HG002", neighbor_threshold="0.4,0.6", ins_threshold=0.4,
                del_threshold=0.6, enable_whatshap=False, impute_indel_phase=False),
    "short_ont": dict(sequencing="short_ont", snp_model="ONT-HG002", indel_model="ONT-HG002", neighbor_threshold="0.3,0.7",
                      ins_threshold=0.4, del_threshold=0.6, enable_whatshap=False, impute_indel_phase=False),
    "ul_ont": dict(sequencing="ul_ont", snp_model="ONT-HG002", indel_model="ONT-HG002", neighbor_threshold="0.4,0.6", ins_threshold=0.4,
                   del_threshold=0.6, enable_whatshap=False, impute_indel_phase=False),
    "ul_ont_extreme": dict(sequencing="ul_ont_extreme", snp_model="ONT-HG002", indel_model="ONT-HG002", neighbor_threshold="0.4,0.6",
                           ins_threshold=0.4, del_threshold=0.6, enable_whatshap=False, impute_indel_phase=False),
    "ccs": dict(sequencing="pacbio", snp_model="CCS-HG002", indel_model="CCS-HG002", neighbor_threshold="0.3,0.7", ins_threshold=0.4,
                del_threshold=0.4, enable_whatshap=True, impute_indel_phase=True),
    "clr": dict(sequencing="pacbio", snp_model="CLR-HG002", indel_model="ONT-HG002", neighbor_threshold="0.3,0.6", ins_threshold=0.6,
                del_threshold=0.6, win_size=10, small_win_size=2, enable_whatshap=True, impute_indel_phase=False),
}


def build_parser():
    p = argparse.ArgumentParser(prog="nanocaller_b200", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    req = p.add_argument_group("Required Arguments")
    req.add_argument("--bam", required=True, help="BAM file (haplotagged with HP/PS tags for indel calling)")
    req.add_argument("--ref", required=True, help="reference FASTA")
    p.add_argument("--preset", choices=["ont", "ul_ont", "ul_ont_extreme", "ccs", "clr"], default=None)
    p.add_argument("--mode", choices=["snps", "indels", "all"], default="all")
    p.add_argument("--sequencing", choices=["short_ont", "ont", "ul_ont", "ul_ont_extreme", "pacbio"], default="ont")
    p.add_argument("--cpu", type=int, default=1, help="sets the chunk grid exactly as in the reference (utils.py:72)")
    p.add_argument("--mincov", type=int, default=4)
    p.add_argument("--maxcov", type=int, default=160)
    p.add_argument("--suppress_progress_bar", action="store_true", default=False)
    p.add_argument("--haploid_genome", action="store_true", default=False)
    p.add_argument("--haploid_X", action="store_true", default=False)
    p.add_argument("--verbose", action="store_true", default=False)
    p.add_argument("--output", type=str, default=None)
    p.add_argument("--prefix", type=str, default="variant_calls")
    p.add_argument("--sample", type=str, default="SAMPLE")
    p.add_argument("--regions", nargs="*")
    p.add_argument("--bed", type=str, default=None)
    p.add_argument("--wgs_contigs", choices=["chr1-22XY", "1-22XY"], default=None)
    p.add_argument("--exclude_bed", type=str, default=None, help="hg38 | hg19 | mm10 | mm39 (needs --nanocaller_src) or a BED path")
    p.add_argument("--snp_model", default="ONT-HG002")
    p.add_argument("--min_allele_freq", type=float, default=0.15)
    p.add_argument("--min_nbr_sites", type=int, default=1)
    p.add_argument("--neighbor_threshold", type=str, default="0.4,0.6")
    p.add_argument("--supplementary", action="store_true", default=False)
    p.add_argument("--disable_coverage_normalization", action="store_true", default=False)
    p.add_argument("--indel_model", default="ONT-HG002")
    p.add_argument("--ins_threshold", type=float, default=0.4)
    p.add_argument("--del_threshold", type=float, default=0.6)
    p.add_argument("--win_size", type=int, default=40)
    p.add_argument("--small_win_size", type=int, default=4)
    p.add_argument("--impute_indel_phase", action="store_true", default=False)
    p.add_argument("--write_phased_bam", action="store_true", default=False,
                   help="also write intermediate_phase_files/{contig}.phased.bam (indelCaller.py:244): the contig's records copied whole with "
                        "the HP / PS tags of the phasing step. The indel stage does not need the file (the tags are staged from memory).")
    p.add_argument("--host_bam_reader", action="store_true", default=False,
                   help="inflate and decode the BAM on the host (libnc_bamio, zlib) instead of on the GPU. The device reader is the default; it "
                        "needs the inflated file to fit in device memory and falls back to the host reader by itself when it does not.")
    p.add_argument("--decompose_indels", action="store_true", default=False,
                   help="normalise the indel records like the reference's `rtg vcfdecompose | rtg vcffilter --non-snps-only` step "
                        "(indelCaller.py:391) with this package's own rule set (host/vcf_decompose.py); the records as the indel stage "
                        "wrote them are kept in intermediate_indel_files/{prefix}.raw.indel.vcf. Off by default: not pinned against rtg.")
    p.add_argument("--phase", action="store_true", default=False)
    p.add_argument("--phase_qual_score", type=float, default=10)
    p.add_argument("--enable_whatshap", action="store_true", default=False)
    own = p.add_argument_group("nanocaller_b200")
    own.add_argument("--device", type=int, default=0, help="CUDA device")
    own.add_argument("--nanocaller_src", type=str, default=None, help="a NanoCaller checkout, for models / BED files that are not bundled")
    return p


def parse_args(argv):
    """argparse + the preset rule of NanoCaller:164-174: a preset value applies unless the flag was given explicitly."""
    args = build_parser().parse_args(argv)
    args.supplementary = False                                   # NanoCaller:160
    set_flags = {x.replace("-", "").split("=")[0] for x in argv if x.startswith("--")}
    if args.preset:
        for k, v in PRESETS[args.preset].items():
            if k not in set_flags:
                setattr(args, k, v)
    if not args.output:
        args.output = os.getcwd()
    return args


def get_regions_list(args, contig_lengths):
    """utils.get_regions_list (utils.py:6-65).  `contig_lengths`: ordered {name: length} from the BAM header."""
    ploidy = "haploid" if args.haploid_genome else "diploid"
    regions = []
    if args.wgs_contigs:
        for c in list(range(1, 23)) + ["X", "Y"]:
            name = ("chr%s" % c) if args.wgs_contigs == "chr1-22XY" else str(c)
            if name in contig_lengths:
                regions.append([name, 1, contig_lengths[name], ploidy])
    elif args.regions:
        for r in args.regions:
            r2 = r.split(":")
            if len(r2) == 1:
                if r2[0] in contig_lengths:
                    regions.append([r2[0], 1, contig_lengths[r2[0]], ploidy])
                else:
                    print("\n%s: Contig %s not present in the BAM file." % (datetime.datetime.now(), r2[0]), flush=True)
            elif len(r2) == 2 and len(r2[1].split("-")) == 2:
                a, b = r2[1].split("-")
                regions.append([r2[0], int(a), int(b), ploidy])
            else:
                print("\n%s: Invalid region %s." % (datetime.datetime.now(), r), flush=True)
    elif args.bed:
        with open(args.bed) as f:
            for line in f:
                t = line.rstrip("\n").split()
                if not t:
                    continue
                if t[0] in contig_lengths:
                    regions.append([t[0], int(t[1]), int(t[2]), ploidy])
                else:
                    print("\n%s: Contig %s not present in the BAM file." % (datetime.datetime.now(), t[0]), flush=True)
    else:
        regions = [[c, 1, n, ploidy] for c, n in contig_lengths.items()]
    if not regions:
        print("\n%s: No valid regions found." % datetime.datetime.now(), flush=True)
        raise SystemExit(2)
    for r in regions:                                             # utils.py:54-60
        if r[0] in ("chrY", "Y", "chrM", "M"):
            r[3] = "haploid"
        elif r[0] in ("chrX", "X"):
            r[3] = "haploid" if args.haploid_X else "diploid"
    return [tuple(r) for r in regions]


def get_chunks(regions_list, cpu, max_chunk_size=500000, min_chunk_size=10000, total=None):
    """utils.py:67-83: inclusive chunk ends, shared between neighbours.  `total`: the base count that sizes the chunks when
    `regions_list` is one rank's share of a multi-GPU run (host/multi.py) — the reference sums over ALL regions (:72)."""
    if total is None:
        total = sum(r[2] - r[1] + 1 for r in regions_list)
    size = min(max_chunk_size, max(min_chunk_size, total // cpu + 1))
    return [{"chrom": c, "start": s, "end": min(end, s + size), "ploidy": pl}
            for c, start, end, pl in regions_list for s in range(start, end, size)]


def _load_exclude_bed(args):
    """-> registered BED key or None.  Accepts the reference's genome names (files live in a NanoCaller checkout) or a path."""
    from .host import sources
    name = args.exclude_bed
    if not name:
        return None
    path = name
    if name in ("hg38", "hg19", "mm10", "mm39"):
        if not args.nanocaller_src:
            raise SystemExit("--exclude_bed %s: the BED files are not bundled; pass --nanocaller_src <NanoCaller checkout>" % name)
        path = os.path.join(args.nanocaller_src, "nanocaller_src/release_data/bed_files/%s_centro_telo.bed.gz" % name)
    opener = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    table = {}
    with opener(path, "rt") as f:
        for line in f:
            t = line.split()
            if len(t) >= 3 and not line.startswith(("#", "track", "browser")):
                table.setdefault(t[0], []).append((int(t[1]), int(t[2])))
    sources.register_bed(path, table)
    return path


def _contig_lengths(bam):
    from .host import bamio, sources
    return sources.contig_lengths(bam) if bam.startswith("mem://") else bamio.bam_contigs(bam)


def _groups(chunks):
    """Consecutive chunks of one (contig, ploidy): one staging + one batched launch sequence each."""
    out = []
    for c in chunks:
        if out and out[-1][0]["chrom"] == c["chrom"] and out[-1][0]["ploidy"] == c["ploidy"]:
            out[-1].append(c)
        else:
            out.append([c])
    return out


def _phase_stage(args, regions, chrom_list, out):
    """indelCaller.phase_run (indelCaller.py:190-262): PASS SNP records of every diploid contig -> phased records + haplotagged
    reads (host/phasing.py); haploid contigs pass through (:191-199).  Writes `{prefix}.snps.phased.vcf.gz` (:360)."""
    from .host import phasing, sources, vcfio
    t1 = time.time()
    if getattr(args, "enable_whatshap", False):
        print("\n%s: note: --enable_whatshap (WhatsHap's --distrust-genotypes --include-homozygous, indelCaller.py:225) is not reproduced: "
              "genotypes are phased as called." % datetime.datetime.now(), flush=True)
    by_chrom = {}
    for ln in vcfio.read_records(out["snps"]):
        by_chrom.setdefault(ln.split("\t", 1)[0], []).append(ln)
    ploidy_of = {}
    for r in regions:
        ploidy_of.setdefault(r[0], r[3])
    phased_lines, pstats = [], {}
    for chrom in chrom_list:
        lines_c = by_chrom.get(chrom, [])
        if ploidy_of[chrom] == "haploid":
            phased_lines += lines_c
            continue
        rs = sources.resolve(args.bam, chrom)
        tagged = rs.n_tagged if isinstance(rs, sources.DeviceContig) else int((rs.hp > 0).sum())
        if not args.phase and tagged > 0:
            print("\n%s: %s: reads are already haplotagged; HP/PS tags of the BAM are used (give --phase to re-phase)."
                  % (datetime.datetime.now(), chrom), flush=True)
            phased_lines += lines_c
            continue
        rs = sources.host_reads(args.bam, chrom)                  # phasing walks the reads on the host
        new_lines, st = phasing.phase_snp_records(lines_c, rs, args.phase_qual_score, supplementary=args.supplementary)
        phased_lines += new_lines
        pstats[chrom] = st
        if args.write_phased_bam and os.path.exists(args.bam):
            from .host import bamio
            pdir = os.path.join(args.output, "intermediate_phase_files")
            os.makedirs(pdir, exist_ok=True)
            st["phased_bam"] = bamio.write_haplotagged_bam(args.bam, chrom, rs.hp, rs.ps, os.path.join(pdir, "%s.phased.bam" % chrom))
    partial = bool(getattr(args, "_partial", False))              # one rank of a multi-GPU run: plain text, merged and compressed by rank 0
    php = os.path.join(args.output, "%s.snps.phased.vcf%s" % (args.prefix, "" if partial else ".gz"))
    vcfio.write_vcf(php, "phased_snps", chrom_list, phased_lines, args.sample, index=not partial)
    out.update(phased_snps=php, phase_stats=pstats, phase_seconds=time.time() - t1)
    print("\n%s: Phasing completed. Time taken= %.4f\n" % (datetime.datetime.now(), time.time() - t1), flush=True)


def run(args):
    from .host import bamio, indel_caller, models, snp_caller, snp_pileups, vcfio
    t0 = time.time()
    threshold = [float(x) for x in args.neighbor_threshold.split(",")[:2]]
    os.makedirs(args.output, exist_ok=True)
    with open(os.path.join(args.output, "args"), "w") as f:      # NanoCaller:182-186
        f.write("Command: python %s\n\n\n" % " ".join(sys.argv))
        f.write("------Parameters Used For Variant Calling------\n")
        for k, v in vars(args).items():
            f.write("{}: {}\n".format(k, v))
    contig_lengths = _contig_lengths(args.bam)                    # utils.py:9-50 asks the BAM header, not the FASTA
    regions = getattr(args, "_regions", None) or get_regions_list(args, contig_lengths)     # _regions: one rank's share (host/multi.py)
    ctx = snp_pileups.context(args.device)                        # fails loudly without an sm_100 device
    if not args.bam.startswith("mem://"):                         # mem://: an in-memory source registered by the caller (host/sources.py)
        wanted = {r[0] for r in regions}                          # only the contigs that will be called
        opened = False
        if not args.host_bam_reader and not getattr(args, "_read_windows", None):
            from .host import capi
            try:                                                  # BGZF inflate + record decoding on the GPU: only compressed bytes cross PCIe
                bamio.open_alignment_device(ctx, args.bam, args.ref, contigs=wanted)
                opened = True
            except capi.NcError as e:
                print("\n%s: device BAM reader not used (%s); reading on the host." % (datetime.datetime.now(), e), flush=True)
        if not opened:
            bamio.open_alignment(args.bam, args.ref, contigs=wanted)
    if getattr(args, "_read_windows", None):                      # one rank of a chunk-sharded run (host/multi.py): keep its part of every contig
        from .host import sources
        sources.restrict(args.bam, args._read_windows)
    exclude = _load_exclude_bed(args)
    chrom_list = list(dict.fromkeys(r[0] for r in regions))
    # one rank of a multi-GPU run (host/multi.py): its record files are read back by the merge on rank 0, so they stay plain text
    # without an index; the reference's compressed, indexed outputs are written once, by rank 0
    partial = bool(getattr(args, "_partial", False))
    gz, idx = ("" if partial else ".gz"), not partial
    ctx = snp_pileups.context(args.device)                        # fails loudly without an sm_100 device
    out = {}
    t_read = time.time() - t0

    if args.mode in ("snps", "all"):
        t1 = time.time()
        params = dict(sam_path=args.bam, fasta_path=args.ref, mincov=args.mincov, maxcov=args.maxcov, min_allele_freq=args.min_allele_freq,
                      min_nbr_sites=args.min_nbr_sites, threshold=threshold, seq=args.sequencing, supplementary=args.supplementary,
                      exclude_bed=exclude, disable_coverage_normalization=args.disable_coverage_normalization)
        tensors, cov = models.get_SNP_model(args.snp_model, args.nanocaller_src)
        if tensors is None:
            print("Invalid SNP model name or path", flush=True)     # snpCaller.py:66-68
            raise SystemExit(2)
        chunks = get_chunks(regions, args.cpu, total=getattr(args, "_total_bases", None))
        hap = None
        if any(c["ploidy"] == "haploid" for c in chunks):
            hap = models.get_SNP_model("haploid", args.nanocaller_src)[0]          # snpCaller.py:74
        parts = []
        for grp in _groups(chunks):
            blob, off, ok, pos = snp_caller.call_chunks_blob(params, grp, (tensors, cov), hap_weights=hap, device=args.device)
            parts.append((grp[0]["chrom"], blob, off, ok, pos))
        allp = os.path.join(args.output, "%s.unfiltered.snps.vcf%s" % (args.prefix, gz))
        passp = os.path.join(args.output, "%s.snps.vcf%s" % (args.prefix, gz))
        n_all = vcfio.write_vcf_blobs(allp, "snps", chrom_list, parts, args.sample, index=idx)      # + .csi, like tabix -fp vcf --csi (snpCaller.py:283)
        vcfio.write_vcf_blobs(passp, "snps", chrom_list, parts, args.sample, pass_only=True, index=idx)
        out.update(unfiltered_snps=allp, snps=passp, n_snp_records=n_all, snp_seconds=time.time() - t1)
        print("\n%s: SNP calling completed. Time taken= %.4f\n" % (datetime.datetime.now(), time.time() - t1), flush=True)

    if args.mode == "all" or (args.mode == "snps" and args.phase):      # indelCaller.phase_run (indelCaller.py:190-262), NanoCaller:41
        _phase_stage(args, regions, chrom_list, out)

    if args.mode in ("indels", "all"):
        t1 = time.time()
        params = dict(sam_path=args.bam, fasta_path=args.ref, mincov=args.mincov, maxcov=args.maxcov, seq=args.sequencing,
                      del_t=args.del_threshold, ins_t=args.ins_threshold, impute_indel_phase=args.impute_indel_phase,
                      supplementary=args.supplementary, exclude_bed=exclude, win_size=args.win_size, small_win_size=args.small_win_size)
        ind = models.get_indel_model(args.indel_model, args.nanocaller_src)
        if ind is None:
            print("Invalid indel model name or path", flush=True)
            raise SystemExit(2)
        chunks = get_chunks(regions, args.cpu, max_chunk_size=100000, total=getattr(args, "_total_bases", None))
        hap_ind = None
        if any(c["ploidy"] == "haploid" for c in chunks):
            hap_ind = models.get_indel_model("haploid", args.nanocaller_src)
        lines = []
        for grp in _groups(chunks):                               # indelCaller.py:327-336 hands every chunk its (phased) BAM
            lines += indel_caller.call_chunks(params, [dict(c, sam_path=args.bam) for c in grp], ind, hap_tensors=hap_ind, device=args.device)
        indp = os.path.join(args.output, "%s.indels.vcf%s" % (args.prefix, gz))
        if args.decompose_indels:                                 # indelCaller.py:369,:391
            from .host import vcf_decompose
            raw_dir = os.path.join(args.output, "intermediate_indel_files")
            os.makedirs(raw_dir, exist_ok=True)
            out["raw_indels"] = os.path.join(raw_dir, "%s.raw.indel.vcf" % args.prefix)
            vcfio.write_vcf(out["raw_indels"], "indels", chrom_list, lines, args.sample)
            lines = vcf_decompose.decompose_records(vcfio.sort_records(lines, chrom_list), contigs=chrom_list)
        vcfio.write_vcf(indp, "indels", chrom_list, lines, args.sample, index=idx)
        out.update(indels=indp, n_indel_records=len(lines), indel_seconds=time.time() - t1)
        print("\n%s: Indel calling completed. Time taken= %.4f\n" % (datetime.datetime.now(), time.time() - t1), flush=True)
        if args.mode == "all":
            final = os.path.join(args.output, "%s.vcf%s" % (args.prefix, gz))
            snp_lines = vcfio.read_records(out["phased_snps"])             # indelCaller.py:395 concatenates the phased SNP file
            vcfio.write_vcf(final, "all", chrom_list, snp_lines + lines, args.sample, index=idx)
            out["final"] = final
    out.update(read_seconds=t_read, launches=ctx.timings()["launches"])
    return out


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    t = time.time()
    args = parse_args(argv)
    if args.regions and args.bed:
        print("\n%s: Please use either --regions or --bed but not both." % datetime.datetime.now(), flush=True)
        raise SystemExit(2)
    print("\n%s: Starting nanocaller_b200.\n" % datetime.datetime.now(), flush=True)
    from .host import multi
    rank, world, local = multi.env_world()
    if world > 1:                                                 # torchrun: one process per GPU, contigs sharded (host/multi.py)
        import torch
        import torch.distributed as dist
        from .host import bamio
        cuda = torch.cuda.is_available()
        if cuda:
            torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl" if cuda else "gloo")
        regions = get_regions_list(args, _contig_lengths(args.bam))
        out = multi.run_distributed(args, run, regions, dist, device=("cuda:%d" % local) if cuda else "cpu")
        print("\n%s: Total Time Elapsed: %.2f seconds" % (datetime.datetime.now(), time.time() - t))
        return out
    out = run(args)
    print("\n%s: Total Time Elapsed: %.2f seconds" % (datetime.datetime.now(), time.time() - t))
    return out


if __name__ == "__main__":
    main()
