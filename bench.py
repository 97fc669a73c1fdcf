#!/usr/bin/env python
"""bench.py — candidate sites/sec of the SNP hot path (pileup scan + tensor build + CNN forward).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the oracle port on all host cores

Workload (BASELINE.json configs[1]): mode=snps, ONT preset, ONT-HG002 snp_model, synthetic chr20-shape
60 Mb contig at 30x, reference chunk grid (120 chunks of 500 kb, `--cpu 1`).  With N > 1 every rank owns one
such contig (different seed): chunks are independent units, so ranks share nothing on the data path and the
only collective is the gather of per-site call records to rank 0 (weak scaling).

A step = one pass of the hot path over the rank's contig:
  value  device-resident: BAM-native arrays already in HBM; K0 decode + K1 scan + K2 tensors + CNN, timed with
         CUDA events on the library's stream.
  e2e    through the C-ABI with HOST (pinned) buffers: nc_stage_reads (H2D) + the same kernels + D2H of the per-site
         call records (probabilities + site metadata) (+ NCCL gather when N > 1), timed on the host around
         synchronised steps.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SITE = 3_455_760          # 2 x 1,727,880 MAC, dense-equivalent
DCT = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont",
           supplementary=False, exclude_bed=None)
MODEL = "ONT-HG002"


def workload(rank, length):
    from nanocaller_b200.synth import make_world
    from nanocaller_b200.cli import get_chunks     # the product's chunk grid (utils.py:67-83); oracle/ is only used by the CPU arm
    rs = make_world(chrom="chr20", preset="ont", contig_len=length, seed=20 + rank, coverage=30.0).reads
    chunks = get_chunks([("chr20", 1, length, "diploid")], 1)
    return rs, chunks


def measured_traffic():
    """DRAM bytes per site of the kernel groups, from the committed `ncu --set full` capture (profiles/r1_kernel_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum per launch / sites of that launch).  None when the file is absent."""
    p = os.path.join(ROOT, "profiles", "r1_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "samples": len(sm), "reasons": sorted(reasons)}


def pinned_copy(a):
    import torch
    t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype)) if str(a.dtype) != "uint16" else torch.int16, pin_memory=True)
    v = t.numpy().view(a.dtype)
    v[...] = a
    return t, v


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_chunk(args):
    """One chunk through the CPU restatement: pileup tensors + coverage scaling + fp32 CNN (1 thread)."""
    import torch
    torch.set_num_threads(1)
    from oracle import cnn_oracle, snp_oracle
    rs, tensors, tc, chunk = args if len(args) == 4 else (_G["rs"], _G["tensors"], _G["tc"], args[0])
    pos, ref, mat, dp, freq, depth, fwd, rev = snp_oracle.get_snp_testing_candidates(rs, DCT, chunk)
    if len(pos) == 0:
        return 0
    x = snp_oracle.scale_counts(mat, tc, coverage=float(depth))
    for b in range(0, len(x), 1000):                                   # batch_size=1000, snpCaller.py:80
        cnn_oracle.snp_probs(tensors, x[b:b + 1000], np.asarray(ref[b:b + 1000], np.float32))
    return len(pos)


_G = {}


def cpu_sample_chunks(length, n, size):
    """n sub-chunks of `size` bp spread evenly over the contig (a bounded sample of the same workload)."""
    starts = np.linspace(100_000, max(100_001, length - size - 100_000), n).astype(int)
    return [{"chrom": "chr20", "start": int(s) + 1, "end": int(s) + size, "ploidy": "diploid"} for s in starts]


def cpu_pool_run(pool, chunks):
    t = time.perf_counter()
    sites = sum(pool.map(_cpu_chunk, [(c,) for c in chunks], chunksize=1))
    return sites, time.perf_counter() - t


def make_pool(rs, cores):
    import multiprocessing as mp
    from nanocaller_b200.host import weights as W
    tensors, meta = W.load_model("snp", MODEL)
    _G.update(rs=rs, tensors=tensors, tc=meta["train_coverage"])
    return mp.get_context("fork").Pool(cores)      # fork: workers share the read set copy-on-write, like mp.Process in snpCaller.py:238


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU restatement (oracle/, kind "port": the unmodified reference
    needs pysam/TensorFlow, absent here and on the GPU box) on every host core, one process per core pulling chunks,
    mirroring snpCaller.call_manager (snpCaller.py:213-245)."""
    if rank != 0:
        return
    length = int(os.environ.get("NC_BENCH_LEN", 60_000_000))
    cores = os.cpu_count() or 1
    rs, _ = workload(0, length)
    total = max(1, args.steps + args.warmup)
    size = int(min(500_000, max(50_000, (150.0 / total / 2.5) * 100_000)))   # ~2.5 s of oracle work per 100 kb
    size = min(size, max(10_000, length // 4))
    pool = make_pool(rs, cores)
    chunks = cpu_sample_chunks(length, cores, size)
    for _ in range(args.warmup):
        cpu_pool_run(pool, chunks)
    sites = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s, _ = cpu_pool_run(pool, chunks)
        sites += s
    dt = time.perf_counter() - t0
    pool.close()
    v = sites / dt if dt > 0 else 0.0
    sample = "%d sub-chunks of %d bp per step (one per core) of the 60 Mb contig, numpy/torch-CPU restatement (oracle/)" % (cores, size)
    print(json.dumps({
        "impl": "reference", "metric": "candidate sites/sec (pileup+CNN)", "value": v, "unit": "sites/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "mode=snps, ONT preset, %s snp_model, synthetic chr20-shape %d bp @30x" % (MODEL, length), "sample": sample},
        "cpu_baseline": {"value": v, "unit": "sites/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cnn-impl", type=int, default=0, help="0 tcgen05 (default), 1 fp32 CUDA cores")
    ap.add_argument("--from-bam", type=int, default=0, metavar="BP",
                    help="also report sites/s from a BAM + FASTA on disk (SURVEY 8d figure ii) for a synthetic contig of BP bases: "
                         "native BGZF inflate + record copy into pinned memory, H2D, kernels, D2H (off by default: writing the BAM takes a while)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.warmup = max(args.warmup, 3)

    from nanocaller_b200.host import capi, weights as W
    from nanocaller_b200.host.gather import gather_calls
    length = int(os.environ.get("NC_BENCH_LEN", 60_000_000))
    rs, chunks = workload(rank, length)
    ch = [(c["start"], c["end"]) for c in chunks]
    ctx = capi.Context(local)
    tensors, meta = W.load_model("snp", MODEL)
    ctx.load_snp_weights(W.pack_snp_blob(tensors, False), meta["train_coverage"], False)
    params = capi.snp_params(DCT, "diploid")

    # host buffers in pinned memory (the BAM-native arrays a reader would hand over)
    keep, arrs = [], []
    for a in (rs.pos, rs.flag, rs.cigar_off, rs.cigar, rs.seq_off, rs.l_seq, rs.seq4, rs.ref):
        t, v = pinned_copy(a)
        keep.append(t); arrs.append(v)
    h2d = int(sum(a.nbytes for a in arrs))

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.sync(); torch.cuda.synchronize()

    def step_resident():
        ctx.invalidate_decode()
        n = ctx.snp_scan(params, ch)
        ctx.snp_forward(normalize=True, impl=args.cnn_impl, fetch=False)
        return n

    pin_probs = pin_meta = None

    def step_e2e():
        nonlocal pin_probs, pin_meta
        ctx.stage_arrays(*arrs)
        n = ctx.snp_scan(params, ch)
        ctx.snp_forward(normalize=True, impl=args.cnn_impl, fetch=False)
        if world > 1:
            return n, gather_calls(ctx, dist, rank, world)
        if pin_probs is None or len(pin_probs) < n:
            tp = torch.empty((int(n * 1.2) + 16, 4), dtype=torch.float32, pin_memory=True)
            tm = torch.empty((int(n * 1.2) + 16, capi.META_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
            keep.extend([tp, tm]); pin_probs, pin_meta = tp.numpy(), tm.numpy()
        ctx.fetch_calls(pin_probs[:n], pin_meta[:n])
        return n, n

    # ---- device-resident figure
    ctx.stage_arrays(*arrs)
    for _ in range(args.warmup):
        n_sites = step_resident()
    l0 = ctx.timings()["launches"]
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ctx.event_record(0)
    acc = {"decode_ms": 0.0, "scan_ms": 0.0, "tensor_ms": 0.0, "cnn_ms": 0.0, "cnn_a_ms": 0.0}
    for _ in range(args.steps):
        n_sites = step_resident()
        tm = ctx.timings()
        for k in acc:
            acc[k] += tm[k]
    ctx.event_record(1)
    dev_ms = ctx.event_elapsed_ms(0, 1)
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = ctx.timings()["launches"] - l0
    tbytes = ctx.timings()["tensor_bytes"]
    # SURVEY 8(d) algorithmic bytes of the tensor build: 5 B per (sampled read, real column) + 1025 int16 + 64 B of site data,
    # from the site records of the last step (untimed)
    survey_bytes = None
    try:
        import numpy as _np
        _m = _np.empty(max(1, int(n_sites)), capi.META_DTYPE)
        _pr = _np.empty((max(1, int(n_sites)), 4), _np.float32)
        ctx.fetch_calls(_pr, _m)
        _m = _m[:int(n_sites)]
        _cols = _m["n_left"].astype(_np.int64) + _m["n_right"].astype(_np.int64) + 1
        survey_bytes = int((5 * _m["sample_depth"].astype(_np.int64) * _cols).sum() + int(n_sites) * (1025 * 2 + 64))
    except Exception as _e:                                   # a reporting extra must never cost the bench line
        sys.stderr.write("bench: survey byte count skipped: %r\n" % (_e,))

    # ---- end-to-end figure, serial: stage -> kernels -> fetch, one contig after the other
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_e2e, n_gathered = step_e2e()
    ctx.sync()
    e2e_serial_s = time.perf_counter() - t0
    barrier()

    # ---- end-to-end figure, pipelined: two contexts (two streams) per GPU, each looping over whole steps, so the H2D copy
    #      of one contig overlaps the kernels of the other — how a run over many contigs is driven.  Same work per step,
    #      same API calls.  With N > 1 every step still ends with the gather of its call records to rank 0; the two host
    #      threads take turns (step order) so that all ranks issue the collectives in the same order.
    e2e_s = e2e_serial_s
    pipelined = False
    if args.steps >= 2:
        import threading
        ctx2 = capi.Context(local)
        ctx2.load_snp_weights(W.pack_snp_blob(tensors, False), meta["train_coverage"], False)
        ctxs = [ctx, ctx2]
        blocking = world >= 4                           # 2 x N host threads on one host: sleeping waits beat spinning ones from 4 ranks on
        for c_ in ctxs:
            c_.set_blocking_sync(blocking)
        bufs = []
        for c in ctxs:
            tp = torch.empty((int(n_sites * 1.2) + 16, 4), dtype=torch.float32, pin_memory=True)
            tm2 = torch.empty((int(n_sites * 1.2) + 16, capi.META_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
            keep.extend([tp, tm2]); bufs.append((tp.numpy(), tm2.numpy()))
        cv = threading.Condition()
        turn = [0]
        gathered = [0]

        def worker(i, steps_of_i, out):
            torch.cuda.set_device(local)                   # the current device is per host thread
            c, (pp, pm) = ctxs[i], bufs[i]
            for sidx in steps_of_i:
                c.stage_arrays(*arrs)
                n = c.snp_scan(params, ch)
                c.snp_forward(normalize=True, impl=args.cnn_impl, fetch=False)
                if world > 1:
                    with cv:
                        cv.wait_for(lambda: turn[0] == sidx)
                    gathered[0] = gather_calls(c, dist, rank, world)
                    with cv:
                        turn[0] += 1
                        cv.notify_all()
                else:
                    c.fetch_calls(pp[:n], pm[:n])
                    gathered[0] = n
                out[i] += n

        def run(nsteps_total):
            out = [0, 0]
            turn[0] = 0
            th = [threading.Thread(target=worker, args=(i, list(range(i, nsteps_total, 2)), out)) for i in range(2)]
            barrier()
            t = time.perf_counter()
            for x in th:
                x.start()
            for x in th:
                x.join()
            for c in ctxs:
                c.sync()
            dt = time.perf_counter() - t
            barrier()
            return sum(out), dt

        run(max(2, args.warmup))
        sites_p, e2e_p = run(args.steps)
        if sites_p == n_sites * args.steps:
            e2e_s, pipelined = e2e_p, True
            n_gathered = gathered[0]
        ctx2.close()

    from_bam = None
    if args.from_bam > 0 and world == 1:
        import tempfile
        from nanocaller_b200.host import bamio
        from nanocaller_b200.synth import make_world
        rs_d = make_world(chrom="chr20", preset="ont", contig_len=args.from_bam, seed=20, coverage=30.0).reads
        tmpd = tempfile.mkdtemp(prefix="nc_bench_")
        bam_p, fa_p = os.path.join(tmpd, "d.bam"), os.path.join(tmpd, "d.fa")
        bamio.write_bam(bam_p, [rs_d]); bamio.write_fasta(fa_p, [rs_d])
        from nanocaller_b200.cli import get_chunks as _gc
        ch_d = [(c["start"], c["end"]) for c in _gc([("chr20", 1, args.from_bam, "diploid")], 1)]

        arena = {"buf": None, "off": 0}

        def pinned_alloc(shape, dtype):
            """Bump allocator over one pinned arena, reset every step (a real stager keeps such buffers for the whole run)."""
            nb = int(np.prod(shape)) * np.dtype(dtype).itemsize
            if arena["buf"] is None:
                tot = int((rs_d.pos.nbytes + rs_d.flag.nbytes + rs_d.cigar_off.nbytes + rs_d.cigar.nbytes + rs_d.seq_off.nbytes +
                           rs_d.l_seq.nbytes + rs_d.seq4.nbytes) * 1.05) + (1 << 20)
                tt = torch.empty(tot, dtype=torch.uint8, pin_memory=True)
                keep.append(tt)
                arena["buf"] = tt.numpy()
            o = (arena["off"] + 63) & ~63
            arena["off"] = o + nb
            return arena["buf"][o:o + nb].view(dtype)

        def disk_step():
            arena["off"] = 0
            fasta = bamio.read_fasta(fa_p)
            sets, _ = bamio.read_bam_native(bam_p, fasta, alloc=pinned_alloc)
            r = sets[0]
            ctx.stage_arrays(r.pos, r.flag, r.cigar_off, r.cigar, r.seq_off, r.l_seq, r.seq4, r.ref)
            n = ctx.snp_scan(params, ch_d)
            ctx.snp_forward(normalize=True, impl=args.cnn_impl, fetch=False)
            ctx.fetch_calls(pin_probs[:n], pin_meta[:n])
            return n
        disk_step()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            n_d = disk_step()
        dt = (time.perf_counter() - t0) / reps
        from_bam = {"value": n_d / dt, "unit": "sites/s", "ms_per_step": dt * 1e3, "contig_bp": args.from_bam, "sites": int(n_d),
                    "bam_bytes": os.path.getsize(bam_p), "host_threads": os.cpu_count(),
                    "path": "libnc_bamio (parallel BGZF inflate + record copy into pinned memory) -> nc_stage_reads -> kernels -> D2H; file in the page cache"}
        import shutil
        shutil.rmtree(tmpd, ignore_errors=True)

    tot_sites, dev_ms_max, e2e_max = n_sites, dev_ms, e2e_s
    if world > 1:
        t = torch.tensor([float(n_sites), dev_ms, e2e_s, e2e_serial_s], dtype=torch.float64, device="cuda")
        ts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        tot_sites = int(sum(x[0].item() for x in ts)); dev_ms_max = max(x[1].item() for x in ts); e2e_max = max(x[2].item() for x in ts)
        e2e_serial_s = max(x[3].item() for x in ts)

    if rank == 0:
        hbm, tflops, which = peaks()
        ms_per_step = dev_ms_max / args.steps
        value = tot_sites / (ms_per_step * 1e-3)
        cnn_ms = acc["cnn_ms"] / args.steps
        tensor_ms = acc["tensor_ms"] / args.steps
        ach = FLOP_PER_SITE * n_sites / (cnn_ms * 1e-3) / 1e12
        tr = measured_traffic() or {}
        # operand-read model of the MMA programs (profiles/r1_umma_rate.txt): an M=128, K=16 MMA with both operands in
        # shared memory costs 32 + N/4 cycles whatever the tensor pipe could do; per site: conv1 2 x 1080, conv2 1584, conv3 448, fc1 84
        mma_cycles_per_site = 2 * 1080 + 1584 + 1344 / 3.0 + 27 * 4 * 100 / 128.0
        sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
        mma_floor_ms = mma_cycles_per_site * n_sites / (148 * sm_clock * 1e6) * 1e3
        roof = {"kernel": "CNN forward (tc_trunk_a + tc_trunk_b + tc_fc: tcgen05 kind::f16, fp16 hi/lo split operands, hi*hi + lo*hi + hi*lo)",
                "bound": "tensor", "achieved": ach, "peak": tflops, "unit": "TFLOP/s", "frac": ach / tflops,
                "traffic": (tr["cnn_dram_bytes_per_site"] * n_sites) if "cnn_dram_bytes_per_site" in tr else None,
                "traffic_source": tr.get("source"),
                "algorithmic_bytes": n_sites * (2064 + 2 * 10240 + 2 * 6912 + 16),
                "peak_source": which + ", sustained bf16", "flop_per_site": FLOP_PER_SITE, "ms_per_launch_group": cnn_ms,
                "mma_operand_model": {"cycles_per_site": mma_cycles_per_site, "floor_ms": mma_floor_ms, "frac_of_floor": mma_floor_ms / cnn_ms,
                                      "note": "N <= 128 MMAs are bound by shared-memory operand reads (32 + N/4 cycles each), not by the bf16 peak"}}
        ta_ms = acc["cnn_a_ms"] / args.steps
        if ta_ms > 0:
            ta_flop = 2 * (82_000 + 82_000 + 410_000 + 737_280)                 # conv1_1 + conv1_2 + conv1_3 + conv2 MACs per site
            ta_ach = ta_flop * n_sites / (ta_ms * 1e-3) / 1e12
            ta_floor = (2 * 1080 + 1584) * n_sites / (148 * sm_clock * 1e6) * 1e3
            roof["dominant_kernel"] = {"kernel": "tc_trunk_a_kernel (conv1_1 + conv1_2 + conv1_3 + conv2 of every site; %.0f %% of the step)" % (100 * ta_ms / ms_per_step),
                                       "ms_per_launch": ta_ms, "flop_per_site": ta_flop, "achieved": ta_ach, "peak": tflops, "unit": "TFLOP/s",
                                       "frac": ta_ach / tflops, "mma_operand_floor_ms": ta_floor, "frac_of_operand_floor": ta_floor / ta_ms,
                                       "traffic": (tr.get("per_kernel", {}).get("tc_trunk_a_kernel", {}).get("dram_bytes", 0) / max(1, tr.get("sites", 1)) * n_sites) or None,
                                       "algorithmic_bytes": n_sites * (2064 + 10240)}
        roof2 = {"kernel": "tensor_kernel (K2 pileup tensor build)", "bound": "hbm", "achieved": tbytes / (tensor_ms * 1e-3) / 1e9,
                 "peak": hbm, "unit": "GB/s", "frac": tbytes / (tensor_ms * 1e-3) / 1e9 / hbm,
                 "traffic": (tr["tensor_dram_bytes_per_site"] * n_sites) if "tensor_dram_bytes_per_site" in tr else None,
                 "bytes_per_site": tbytes / max(1, n_sites), "ms_per_launch": tensor_ms,
                 "note": "achieved counts the bytes this kernel has to move (int16 tensor + site record); survey_8d counts SURVEY 8(d)'s "
                         "B2 = 5 B per (sampled read, real column) + 1025 int16 + 64 B, i.e. the read-code lists a per-site kernel would stream"}
        if survey_bytes:
            roof2["survey_8d"] = {"bytes_per_site": survey_bytes / max(1, n_sites), "achieved": survey_bytes / (tensor_ms * 1e-3) / 1e9,
                                  "frac": survey_bytes / (tensor_ms * 1e-3) / 1e9 / hbm}
        cpu = None
        if world == 1:
            cores = os.cpu_count() or 1
            pool = make_pool(rs, cores)
            smp = cpu_sample_chunks(length, cores, min(300_000, max(10_000, length // 4)))
            s, dt = cpu_pool_run(pool, smp)
            pool.close()
            cpu = {"value": s / dt, "unit": "sites/s", "cores": cores, "kind": "port",
                   "sample": "%d sub-chunks of %d bp (one per core), oracle/ numpy+torch-CPU restatement, %.1f s" % (cores, smp[0]["end"] - smp[0]["start"] + 1, dt)}
        d2h = int((n_gathered if world > 1 else n_e2e) * (16 + capi.META_DTYPE.itemsize))      # rank 0 reads the gathered records
        out = {"metric": "candidate sites/sec (pileup+CNN)", "value": value, "unit": "sites/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "int16 pileup counts; CNN fp16 hi/lo split operands with fp32 accumulation" if args.cnn_impl == 0 else "int16 pileup counts; f32 CNN",
               "data": "synthetic",
               "config": {"workload": "configs[1]: mode=snps, ONT preset, %s snp_model, synthetic chr20-shape %d bp @30x per GPU, %d chunks" % (MODEL, length, len(ch)),
                          "sites_per_gpu": n_sites, "aligned_bases_per_gpu": rs.aligned_bases(), "reads_per_gpu": rs.n,
                          "l2": "inputs (%.2f GB) and tensors (%.2f GB) exceed the 126 MB L2, no flush needed" % (h2d / 1e9, n_sites * 2064 / 1e9),
                          "parallelism": "1 rank per GPU, chunks sharded by contig, gather of call records to rank 0" if world > 1 else "single GPU"},
               "phase_ms": {k: v / args.steps for k, v in acc.items()},
               "e2e": {"value": tot_sites / (e2e_max / args.steps), "unit": "sites/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_max / args.steps * 1e3, "gathered_sites": int(n_gathered),
                       "mode": ("2 contexts per GPU, H2D of one contig overlapped with kernels of the other"
                                + ("; every step ends with the NCCL gather of its call records to rank 0" if world > 1 else "")) if pipelined else "serial",
                       "host_waits": "blocking (cudaEventBlockingSync)" if (pipelined and world >= 4) else "spinning",
                       "serial_value": tot_sites / (e2e_serial_s / args.steps)},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_pileup": roof2, "cpu_baseline": cpu}
        if from_bam:
            out["from_bam"] = from_bam
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
