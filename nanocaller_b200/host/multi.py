"""The command line on several GPUs of one box: `python -m torch.distributed.run --nproc-per-node N -m nanocaller_b200 ...`.

The reference fans chunks out to `--cpu` worker processes, each writing its own intermediate VCF, and the parent concatenates
them (snpCaller.py:204-280, indelCaller.py:340-395).  Here one process drives one GPU.  When phasing or indel calling is part
of the run the sharding unit is the CONTIG: phasing between the stages needs all SNP calls and all reads of a contig in one place
(indelCaller.phase_run works per contig as well, indelCaller.py:190); contigs go to ranks longest-first onto the least loaded
rank.  `--mode snps` without `--phase` has no such coupling and is cut into contiguous runs of chunks instead, so one long contig
spreads over all GPUs.  Either way chunks stay what they are on one GPU — the chunk grid is computed from the total of ALL regions
(utils.py:72) — so every record is identical to the single-GPU run.  The one exchange is the gather of the ranks' record text to rank 0 (sizes, then padded byte tensors: NCCL on the
GPU box, gloo in the CPU test), which merges, sorts, compresses and indexes.  (bench.py and host/shard.py shard by chunk.)"""
import copy
import pickle
import os

import numpy as np


def env_world():
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) outside it."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def assign_contigs(regions, world):
    """regions: (contig, start, end, ploidy) tuples -> list of `world` lists of regions; contigs are kept whole, placed longest
    first on the least loaded rank (ties: lowest rank), each rank's regions in input order."""
    span = {}
    for c, s, e, _ in regions:
        span[c] = span.get(c, 0) + (e - s + 1)
    load = [0] * world
    owner = {}
    for c in sorted(span, key=lambda c: (-span[c], list(span).index(c))):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[c] = r
        load[r] += span[c]
    return [[reg for reg in regions if owner[reg[0]] == r] for r in range(world)]


def _contiguous_partition(weights, world):
    """Indices 0..n-1 cut into `world` contiguous runs of about equal weight; every rank gets at least one item when n >= world."""
    n, out, i, rem = len(weights), [], 0, float(sum(weights))
    for r in range(world):
        left = world - r
        if left == 1:
            out.append(list(range(i, n)))
            break
        target, j, acc = rem / left, i, 0.0
        while j < n and (n - j) > (left - 1) and (acc == 0.0 or acc + weights[j] / 2.0 <= target):
            acc += weights[j]
            j += 1
        out.append(list(range(i, j)))
        rem -= acc
        i = j
    return out


def assign_chunk_runs(regions, world, cpu):
    """SNP calling without phasing has no coupling between chunks (snpCaller.py:83-86), so a single long contig can be split as well:
    the chunk grid of the whole run (utils.py:67-83) is cut into `world` contiguous, balanced runs (host/shard.py) and every run is
    handed over as regions that start and end on grid points — `get_chunks(share, cpu, total=whole run)` then reproduces exactly the
    chunks of that run, shared boundaries included.  -> list of `world` lists of (contig, start, end, ploidy)."""
    from ..cli import get_chunks
    chunks = get_chunks(regions, cpu)
    out = []
    for idxs in _contiguous_partition([c["end"] - c["start"] + 1 for c in chunks], world):
        regs = []
        for i in idxs:
            c = chunks[i]
            if regs and regs[-1][0] == c["chrom"] and regs[-1][2] == c["start"] and regs[-1][3] == c["ploidy"]:
                regs[-1][2] = c["end"]
            else:
                regs.append([c["chrom"], c["start"], c["end"], c["ploidy"]])
        out.append([tuple(r) for r in regs])
    return out


READ_MARGIN = 60_000      # the SNP scan piles up [start - 50000, end + 50000] (generate_SNP_pileups.py:156); 10 kb to spare


def read_windows(regions):
    """{contig: (lo0, hi0)} covering every region of the share plus the pileup flank."""
    out = {}
    for c, s, e, _ in regions:
        lo, hi = max(0, s - 1 - READ_MARGIN), e + READ_MARGIN
        out[c] = (min(lo, out[c][0]), max(hi, out[c][1])) if c in out else (lo, hi)
    return out


def gather_bytes(data, dist, rank, world, device="cpu"):
    """Every rank contributes a bytes object; rank 0 gets the list of all of them in rank order, the others None."""
    import torch
    n = torch.tensor([len(data)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=device)
    if len(data):
        buf[:len(data)] = torch.from_numpy(np.frombuffer(data, np.uint8).copy()).to(device)
    if rank == 0:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=0)
        return [p[:s].cpu().numpy().tobytes() for p, s in zip(parts, sizes)]
    dist.gather(buf, None, dst=0)
    return None


OUTPUT_KINDS = (("unfiltered_snps", "%s.unfiltered.snps.vcf.gz", "snps"), ("snps", "%s.snps.vcf.gz", "snps"),
                ("phased_snps", "%s.snps.phased.vcf.gz", "phased_snps"), ("indels", "%s.indels.vcf.gz", "indels"),
                ("final", "%s.vcf.gz", "all"))


def run_distributed(args, run_fn, regions, dist, device="cpu"):
    """Run `run_fn` (cli.run) on this rank's contigs into `{output}/rank{r}` and merge the ranks' records on rank 0 into the
    reference's output names under `{output}`.  -> rank 0: dict like cli.run's; other ranks: {'rank': r}."""
    from . import vcfio
    rank, world, local = env_world()
    by_chunk = getattr(args, "mode", "all") == "snps" and not getattr(args, "phase", False)
    shares = assign_chunk_runs(regions, world, args.cpu) if by_chunk else assign_contigs(regions, world)
    mine = shares[rank]
    chrom_list = list(dict.fromkeys(r[0] for r in regions))
    out_r, failure = {}, None
    if mine:
        sub = copy.copy(args)
        sub.regions = ["%s:%d-%d" % (c, s, e) for c, s, e, _ in mine]      # for the `args` log only
        sub._regions = [tuple(r) for r in mine]                                # the share itself: contig names may hold ':' (HLA-A*01:01:01:01)
        sub.bed = None
        sub.wgs_contigs = None
        sub.output = os.path.join(args.output, "rank%d" % rank)
        sub.device = local
        sub._partial = True                                                   # record files of a rank stay plain text (cli.run)
        sub._total_bases = sum(e - s + 1 for _, s, e, _ in regions)          # utils.py:72 sizes the chunks from ALL regions
        if by_chunk:                                                          # stage only the reads this rank's chunks can see
            sub._read_windows = read_windows(mine)
        try:
            out_r = run_fn(sub)
        except BaseException as e:                                # the other ranks must not wait in a collective for this one
            failure = e
    if dist_any(failure is not None, dist, world, device):
        if failure is not None:
            raise failure
        raise RuntimeError("nanocaller_b200: another rank failed; rank %d stops" % rank)
    merged = {"rank": rank, "world": world, "sharding": "chunk runs" if by_chunk else "contigs",
              "contigs_per_rank": [sorted({r[0] for r in part}) for part in shares], "regions_per_rank": shares}
    # Record order of the reference's outputs = contig order, then position, stable (`bcftools sort`).  A rank's files are already in
    # that order for its share, and the shares are disjoint contigs or consecutive chunk runs, so — as long as the regions of a contig
    # come in ascending order — the merged file is "for every contig, the ranks' blocks in rank order": rank 0 concatenates bytes and
    # builds the CSI index from the (position, REF length, line length) arrays the ranks send along, instead of splitting, sorting
    # and re-parsing millions of lines in Python.
    ascending = _regions_ascending(regions)
    for key, name, kind in OUTPUT_KINDS:
        have = dist_any(key in out_r, dist, world, device)
        if not have:
            continue
        lines = vcfio.read_records(out_r[key]) if key in out_r else []
        if ascending:
            payload = pickle.dumps(_contig_blocks(lines), protocol=pickle.HIGHEST_PROTOCOL)
            parts = gather_bytes(payload, dist, rank, world, device)
            if rank == 0:
                path = os.path.join(args.output, name % args.prefix)
                merged["n_%s_records" % key] = _write_blocks(path, kind, chrom_list, [pickle.loads(p) if p else {} for p in parts], args.sample)
                merged[key] = path
        else:
            parts = gather_bytes("".join(lines).encode(), dist, rank, world, device)
            if rank == 0:
                lines = [ln + "\n" for p in parts for ln in p.decode().split("\n") if ln]
                path = os.path.join(args.output, name % args.prefix)
                vcfio.write_vcf(path, kind, chrom_list, lines, args.sample, index=True)
                merged[key] = path
                merged["n_%s_records" % key] = len(lines)
    dist.barrier()
    return merged


def _regions_ascending(regions):
    """True when, per contig, the regions come in ascending, non-overlapping order (then rank order = genomic order)."""
    last = {}
    for c, s_, e, _ in regions:
        if c in last and s_ < last[c]:
            return False
        last[c] = e
    return True


def _contig_blocks(lines):
    """Record lines of one rank (sorted) -> {contig: (text bytes, pos int64[n], ref allele length int32[n], line length int32[n])}."""
    out = {}
    cur, buf, pos, rl, ll = None, [], [], [], []

    def flush():
        if cur is not None:
            out[cur] = ("".join(buf).encode(), np.asarray(pos, np.int64), np.asarray(rl, np.int32), np.asarray(ll, np.int32))
    for ln in lines:
        f = ln.split("\t", 4)
        if f[0] != cur:
            flush()
            cur, buf, pos, rl, ll = f[0], [], [], [], []
        buf.append(ln); pos.append(int(f[1])); rl.append(len(f[3])); ll.append(len(ln.encode()) if not ln.isascii() else len(ln))
    flush()
    return out


def _write_blocks(path, kind, chrom_list, rank_blocks, sample):
    """header + for every contig the ranks' blocks in rank order -> BGZF + CSI.  -> number of records."""
    from . import vcfio
    head = vcfio.header(kind, chrom_list, sample).encode()
    names = list(chrom_list)
    body, rid, beg, end, nb = [], [], [], [], []
    seen = set(names)
    extra = [c for blocks in rank_blocks for c in blocks if c not in seen and not seen.add(c)]
    for ci, c in enumerate(names + extra):
        for blocks in rank_blocks:
            if c in blocks:
                text, pos, rl, ll = blocks[c]
                body.append(text)
                rid.append(np.full(len(pos), ci, np.int64)); beg.append(pos - 1); end.append(pos - 1 + np.maximum(1, rl)); nb.append(ll.astype(np.int64))
    data = head + b"".join(body)
    if rid:
        rid, beg, end, nb = np.concatenate(rid), np.concatenate(beg), np.concatenate(end), np.concatenate(nb)
        u0 = len(head) + np.concatenate([[0], np.cumsum(nb)[:-1]])
    else:
        rid = beg = end = nb = u0 = np.zeros(0, np.int64)
    vcfio.write_indexed(path, data, records=(names + extra, rid, beg, end, u0, nb))
    return int(len(rid))


def dist_any(flag, dist, world, device="cpu"):
    """True on every rank if `flag` is true on any rank."""
    import torch
    t = torch.tensor([1 if flag else 0], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return int(t.item()) > 0
