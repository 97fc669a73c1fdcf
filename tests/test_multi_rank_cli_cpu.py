"""CPU, world_size 2 over gloo: the command line's multi-GPU driver (host/multi.py) — contigs sharded over ranks, every rank running
the single-GPU flow on its share with the chunk grid of the whole run, record text gathered to rank 0 and merged into the reference's
output names.  The GPU stages are replaced by a deterministic stand-in (`_fake_run`): this checks the plumbing, not the kernels."""
import argparse
import gzip
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from nanocaller_b200 import cli
from nanocaller_b200.host import multi, vcfio

REGIONS = [("chr1", 1, 900_000, "diploid"), ("chr2", 1, 400_000, "diploid"), ("chr3", 1, 350_000, "diploid"),
           ("chrX", 1, 300_000, "diploid"), ("chrY", 1, 100_000, "haploid")]
CONTIGS = [r[0] for r in REGIONS]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_run(args):
    """Stand-in for cli.run: one SNP record per chunk start and one indel record per indel chunk of the regions it is given."""
    os.makedirs(args.output, exist_ok=True)
    regions = []
    for r in args.regions:
        c, se = r.split(":")
        s, e = se.split("-")
        regions.append((c, int(s), int(e), "haploid" if c == "chrY" else "diploid"))
    total = getattr(args, "_total_bases", None)
    snps = ["%s\t%d\t.\tA\tC\t%d.00\tPASS\t.\tGT:DP:VF:AD:ADF:ADR\t0/1:30:0.5000:15,15:8,7:7,8\n" % (ch["chrom"], ch["start"] + 7, 20 + ch["start"] % 9)
            for ch in cli.get_chunks(regions, args.cpu, total=total)]
    indels = ["%s\t%d\t.\tAT\tA\t12.00\tPASS\t.\tGT:GQ\t1/1:9.00\n" % (ch["chrom"], ch["start"] + 3)
              for ch in cli.get_chunks(regions, args.cpu, max_chunk_size=100000, total=total)]
    chrom_list = list(dict.fromkeys(r[0] for r in regions))
    out = {}
    for key, name, kind in multi.OUTPUT_KINDS:
        lines = indels if key == "indels" else (snps + indels if key == "final" else snps)
        out[key] = os.path.join(args.output, name % args.prefix)
        vcfio.write_vcf(out[key], kind, chrom_list, lines, args.sample, index=True)
    return out


def _args(output):
    return argparse.Namespace(regions=None, bed=None, wgs_contigs=None, output=output, prefix="t", sample="S", cpu=3, device=0)


def _worker(rank, world, port, output, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = multi.run_distributed(_args(output), _fake_run, REGIONS, dist)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_command_line_world2_gloo(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path / "multi"), q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _args(str(tmp_path / "single"))
    single.regions = ["%s:%d-%d" % r[:3] for r in REGIONS]
    want = _fake_run(single)
    got = res[0]
    assert got["contigs_per_rank"] == [["chr1", "chrY"], ["chr2", "chr3", "chrX"]] and "snps" not in res[1]
    for key, name, kind in multi.OUTPUT_KINDS:
        assert got[key] == str(tmp_path / "multi" / (name % "t")) and os.path.exists(got[key] + ".csi")
        with gzip.open(got[key], "rt") as f, gzip.open(want[key], "rt") as g:
            assert f.read() == g.read(), key                  # header, contig order, records: identical to the single-process run
    q = vcfio.csi_query(got["final"], "chr3", 0, 200_000)
    assert len(q) > 0 and all(ln.startswith("chr3\t") for ln in q)


def test_assign_contigs_keeps_contigs_whole_and_balances():
    regs = [("a", 1, 1000, "diploid"), ("b", 1, 400, "diploid"), ("a", 2000, 2500, "diploid"), ("c", 1, 700, "diploid"), ("d", 1, 300, "diploid")]
    for world in (1, 2, 3, 8):
        parts = multi.assign_contigs(regs, world)
        assert len(parts) == world and sorted(r for p in parts for r in p) == sorted(regs)
        owners = {}
        for k, p in enumerate(parts):
            for r in p:
                assert owners.setdefault(r[0], k) == k             # a contig lives on one rank
            assert p == [r for r in regs if r in p]               # input order kept
    two = multi.assign_contigs(regs, 2)
    loads = [sum(e - s + 1 for _, s, e, _ in p) for p in two]
    assert loads == [1501, 1400]


def test_chunk_grid_of_a_share_equals_the_grid_of_the_whole_run():
    """utils.py:72 sizes the chunks from the total of all regions: a rank that is handed a subset must be told that total."""
    small = [("c1", 1, 90_000, "diploid"), ("c2", 1, 50_000, "diploid"), ("c3", 1, 30_000, "diploid")]
    whole = cli.get_chunks(small, 4)
    total = sum(e - s + 1 for _, s, e, _ in small)
    parts = multi.assign_contigs(small, 2)
    again = [c for p in parts for c in cli.get_chunks(p, 4, total=total)]
    assert sorted(map(str, again)) == sorted(map(str, whole))
    assert sorted(map(str, [c for p in parts for c in cli.get_chunks(p, 4)])) != sorted(map(str, whole))


def _failing_run(args):
    if any(r.startswith("chr2:") for r in args.regions):
        raise RuntimeError("boom on the rank that owns chr2")
    return _fake_run(args)


def _worker_fail(rank, world, port, output, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        multi.run_distributed(_args(output), _failing_run, REGIONS, dist)
        q.put((rank, "no error"))
    except RuntimeError as e:
        q.put((rank, str(e)))
    dist.destroy_process_group()


def test_a_failing_rank_stops_all_ranks_instead_of_hanging_them(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_fail, args=(r, world, port, str(tmp_path / "f"), q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1].startswith("boom") and "another rank failed" in res[0]


def test_chunk_runs_reproduce_the_grid_of_the_whole_run():
    """`--mode snps`: the ranks' regions start and end on grid points, and their chunk lists put together are the grid of the whole run,
    in order, for small (grid sized by total // cpu) and large (500 kb chunks) inputs."""
    cases = [([("c1", 1, 90_000, "diploid"), ("c2", 1, 50_000, "diploid"), ("c3", 5_000, 30_000, "haploid")], 4),
             ([("chr20", 1, 64_444_167, "diploid")], 16), ([("a", 1, 2_300_000, "diploid"), ("b", 100, 1_200_000, "diploid")], 2)]
    for regs, cpu in cases:
        whole = cli.get_chunks(regs, cpu)
        total = sum(e - s + 1 for _, s, e, _ in regs)
        for world in (1, 2, 3, 8):
            shares = multi.assign_chunk_runs(regs, world, cpu)
            assert len(shares) == world
            again = [c for share in shares for c in cli.get_chunks(share, cpu, total=total)]
            assert again == whole, (regs, cpu, world)
            if len(whole) >= world:
                assert all(shares)                                 # nobody idles when there is a chunk for everyone


def _worker_snps(rank, world, port, output, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = _args(output)
    a.mode, a.phase = "snps", False
    out = multi.run_distributed(a, _fake_run_snps, [("chr1", 1, 900_000, "diploid")], dist)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _fake_run_snps(args):
    out = _fake_run(args)
    return {k: out[k] for k in ("unfiltered_snps", "snps")}


def test_snps_mode_splits_one_contig_over_the_ranks_world2_gloo(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_snps, args=(r, world, port, str(tmp_path / "multi"), q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _args(str(tmp_path / "single"))
    single.regions = ["chr1:1-900000"]
    want = _fake_run_snps(single)
    got = res[0]
    assert got["sharding"] == "chunk runs" and got["contigs_per_rank"] == [["chr1"], ["chr1"]] and "indels" not in got
    r0, r1 = got["regions_per_rank"]
    assert r0[0][1] == 1 and r0[-1][2] == r1[0][1] and r1[-1][2] == 900_000          # the runs meet on a shared grid point
    for key in ("unfiltered_snps", "snps"):
        with gzip.open(got[key], "rt") as f, gzip.open(want[key], "rt") as g:
            assert f.read() == g.read(), key


def test_a_rank_sees_the_same_candidates_through_its_read_window():
    """Chunk-sharded `--mode snps`: a rank stages `ReadSet.window` of its share (+ the 50 kb pileup flank).  The reference's tensors for
    the chunks of that share are the same from the window as from the whole contig (oracle on both)."""
    import numpy as np
    from nanocaller_b200.synth import make_world
    from oracle import snp_oracle
    rs = make_world(chrom="chrW", preset="ont", contig_len=400_000, seed=81, coverage=14.0).reads
    dct = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont", supplementary=False)
    regs = [("chrW", 1, 400_000, "diploid")]
    cpu = 20                                                  # 20,001-base chunks
    shares = multi.assign_chunk_runs(regs, 4, cpu)
    share = shares[2]
    win = multi.read_windows(share)["chrW"]
    sub = rs.window(*win)
    assert 0 < sub.n < 0.75 * rs.n and sub.contig_len == rs.contig_len
    assert sub.checksum() == rs.subset((np.arange(rs.n) >= np.nonzero(rs.ref_end > win[0])[0][0]) & (rs.pos < win[1])).checksum()
    chunks = cli.get_chunks(share, cpu, total=400_000)
    for ch in (chunks[0], chunks[-1]):
        a = snp_oracle.get_snp_testing_candidates(rs, dct, ch)
        b = snp_oracle.get_snp_testing_candidates(sub, dct, ch)
        assert len(a[0]) > 20
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y))
