#!/usr/bin/env python
"""bench.py — candidate sites/sec of the NanoCaller hot path (pileup scan + tensor build + CNN forward).

    python bench.py --gpus N --steps K --warmup W                     # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W    # CPU arm: the reference's own workers on all host cores
    python bench.py --workload all ...                                # BASELINE configs[2] as the main line

Workloads (BASELINE.json):
  snps  configs[1]: mode=snps, ONT preset, ONT-HG002 snp_model, synthetic chr20-shape 60 Mb contig at 30x, reference chunk
        grid (120 chunks of 500 kb, `--cpu 1`).  The default main line.
  all   configs[2]: mode=all (SNP + indel), ONT preset, ONT-HG002 models, synthetic chr1-shape 250 Mb contig at 30x with
        het / hom indels every ~2 kb and HP / PS tags on the reads (the whatshap bypass SURVEY 8(d) states); SNP grid 500
        chunks of 500 kb, indel grid 2,500 chunks of 100 kb.  At N = 1 the default run measures it too and nests the result
        under "mode_all" of the one JSON line (switch off with --no-all).
With N > 1 every rank owns one such contig (different seed): chunks are independent units, ranks share nothing on the data
path, the one collective is the gather of the run's call records to rank 0 (weak scaling).

A step = one pass of the hot path over the rank's contig:
  value  device-resident: BAM-native arrays already in HBM; K0 decode + K1 scan + K2 tensors + CNN (+ the indel scan, slice
         extraction, alignment, tensors and indel CNN for `all`), timed with CUDA events on the library's stream.
  e2e    through the C-ABI with HOST (pinned) buffers: nc_stage_reads (H2D) + the same kernels + D2H of the per-site call
         records (+ for `all` the indel records and consensus strings, and the allele alignment on host threads), timed on
         the host around synchronised steps.
After the timed steps two chunks of the LAST timed step are recomputed by the oracle on the host cores and compared (untimed):
the bench line says what was checked ("parity").  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SITE = 3_455_760          # 2 x 1,727,880 MAC, dense-equivalent
FLOP_PER_INDEL_SITE = 18_946_752   # Indel_model, 2 x 9,473,376 MAC
CNN_BYTES_PER_SITE = 2_050 + 16    # int16 tensor in, four probabilities out
DCT = dict(threshold=[0.4, 0.6], mincov=4, maxcov=160, min_allele_freq=0.15, min_nbr_sites=1, seq="ont",
           supplementary=False, exclude_bed=None)
IDCT = dict(mincov=4, maxcov=160, seq="ont", del_t=0.6, ins_t=0.4, impute_indel_phase=False, supplementary=False,
            win_size=40, small_win_size=4, exclude_bed=None)      # ONT preset, NanoCaller:66
MODEL = "ONT-HG002"
INDEL_CNN_ON_TENSOR_CORES = True       # Indel_model / haploid_Indel_model run on tcgen05 (nc_cnn_tc_indel.cuh) with impl 0
WORKLOADS = {
    "snps": dict(tag="configs[1]", chrom="chr20", length=60_000_000, seed=20, synth={}),
    "all": dict(tag="configs[2]", chrom="chr1", length=250_000_000, seed=1, synth=dict(indel_every=2000, indel_maxlen=50)),
}


def workload_len(kind):
    env = "NC_BENCH_LEN" if kind == "snps" else "NC_BENCH_ALL_LEN"
    return int(os.environ.get(env, WORKLOADS[kind]["length"]))


def workload_name(kind, length):
    """The config string; identical in both arms."""
    if kind == "snps":
        return "configs[1]: mode=snps, ONT preset, %s snp_model, synthetic chr20-shape %d bp @30x per GPU, reference chunk grid (500 kb)" % (MODEL, length)
    return ("configs[2]: mode=all (SNP+indel), ONT preset, %s snp_model + indel_model, synthetic chr1-shape %d bp @30x per GPU, indels every ~2 kb, "
            "HP/PS tags synthetic, reference chunk grids (500 kb SNP, 100 kb indel)" % (MODEL, length))


def make_workload(kind, rank, length):
    from nanocaller_b200.synth import make_world
    from nanocaller_b200.cli import get_chunks     # the product's chunk grid (utils.py:67-83); oracle/ is only used by the CPU arm
    w = WORKLOADS[kind]
    rs = make_world(chrom=w["chrom"], preset="ont", contig_len=length, seed=w["seed"] + rank, coverage=30.0, **w["synth"]).reads
    chunks = get_chunks([(w["chrom"], 1, length, "diploid")], 1)
    ichunks = get_chunks([(w["chrom"], 1, length, "diploid")], 1, max_chunk_size=100000) if kind == "all" else []
    return rs, chunks, ichunks


def measured_traffic():
    """DRAM bytes per site of the kernel groups, from the committed `ncu --set full` capture of this round if present, else last
    round's (dram__bytes_read.sum + dram__bytes_write.sum per launch / sites of that launch).  None when absent."""
    for name in ("r2_kernel_traffic.json", "r1_kernel_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                d = json.load(open(p))
                d.setdefault("file", "profiles/" + name)
                return d
            except Exception:
                pass
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


def bind_to_gpu_numa(local):
    """Run this rank's host threads (and first-touch its pinned buffers) on the CPUs next to its GPU: with 8 ranks the staging
    copies otherwise all read one NUMA node's memory.  -> description string, or None when the topology is not visible."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else str(bdf)).lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]                                    # nvml prints an 8-digit domain, sysfs a 4-digit one
        path = "/sys/bus/pci/devices/%s/local_cpulist" % bdf
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return "cpus %s (PCI %s)" % (txt, bdf)
    except Exception:
        return None
    return None


def pinned_copy(a):
    import torch
    t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype)) if str(a.dtype) != "uint16" else torch.int16, pin_memory=True)
    v = t.numpy().view(a.dtype)
    v[...] = a
    return t, v


# ------------------------------------------------------------------------------------------------ CPU side: oracle port
_G = {}


def _cpu_chunk(args):
    """One SNP (sub-)chunk through the CPU restatement: pileup tensors + coverage scaling + fp32 CNN (1 thread)."""
    import torch
    torch.set_num_threads(1)
    from oracle import cnn_oracle, snp_oracle
    chunk = args[0]
    pos, ref, mat, dp, freq, depth, fwd, rev = snp_oracle.get_snp_testing_candidates(_G["rs"], DCT, chunk)
    if len(pos) == 0:
        return 0
    x = snp_oracle.scale_counts(mat, _G["tc"], coverage=float(depth))
    for b in range(0, len(x), 1000):                                   # batch_size=1000, snpCaller.py:80
        cnn_oracle.snp_probs(_G["tensors"], x[b:b + 1000], np.asarray(ref[b:b + 1000], np.float32))
    return len(pos)


def _cpu_indel_chunk(args):
    """One indel chunk through the CPU restatement: scan, slices, star MSA, tensors, alleles + fp32 indel CNN."""
    import torch
    torch.set_num_threads(1)
    from oracle import cnn_oracle, indel_oracle
    pos, x0, x1, x2, alleles, phase = indel_oracle.get_indel_testing_candidates(_G["rs"], IDCT, args[0])
    if len(pos) == 0:
        return 0
    cnn_oracle.indel_model(_G["itensors"], np.hstack([x0, x1, x2]).astype(np.float32))
    return len(pos)


def _parity_snp_sub(args):
    """Oracle tensors of one sub-range of a chunk (for the in-run parity check)."""
    from oracle import snp_oracle
    pos, ref, mat, dp, freq, depth, fwd, rev = snp_oracle.get_snp_testing_candidates(_G["rs"], DCT, args[0])
    n = len(pos)
    if n == 0:
        return None
    return (np.asarray(pos, np.int64), np.asarray(ref, np.int32), np.asarray(mat).astype(np.int16), np.asarray(dp, np.int64),
            np.asarray(freq, np.float64), float(depth) * n, n, np.asarray(fwd), np.asarray(rev))


def _parity_indel_chunk(args):
    from oracle import cnn_oracle, indel_oracle
    import torch
    torch.set_num_threads(1)
    pos, x0, x1, x2, alleles, phase = indel_oracle.get_indel_testing_candidates(_G["rs"], IDCT, args[0])
    if len(pos) == 0:
        return [], None, None, [], []
    x = np.hstack([x0, x1, x2]).astype(np.float32)
    return list(pos), x, cnn_oracle.indel_model(_G["itensors"], x), alleles, phase


def sample_chunks(chrom, length, n, size):
    """n sub-chunks of `size` bp spread evenly over the contig (a bounded sample of the same workload)."""
    starts = np.linspace(min(100_000, length // 8), max(min(100_000, length // 8) + 1, length - size - min(100_000, length // 8)), n).astype(int)
    return [{"chrom": chrom, "start": int(s) + 1, "end": int(s) + size, "ploidy": "diploid"} for s in starts]


def make_pool(rs, cores, with_indel=False):
    import multiprocessing as mp
    from nanocaller_b200.host import weights as W
    tensors, meta = W.load_model("snp", MODEL)
    _G.update(rs=rs, tensors=tensors, tc=meta["train_coverage"])
    if with_indel:
        _G["itensors"] = W.load_model("indel", MODEL)[0]
    return mp.get_context("fork").Pool(cores)      # fork: workers share the read set copy-on-write, like mp.Process in snpCaller.py:238


# ------------------------------------------------------------------------------------------------ CPU side: the reference itself
def reference_available():
    from oracle import build_ref
    return build_ref.built()


def _count_wrapper(fn, counter):
    """Pass-through wrapper that adds len(pos) of every call to a shared counter (the workers' record files do not list the
    indel candidates that were not called)."""
    def wrapped(*a, **k):
        r = fn(*a, **k)
        with counter.get_lock():
            counter.value += len(r[0])
        return r
    return wrapped


def _ref_snp_worker(params, q, cq, files, counter):
    import torch
    torch.set_num_threads(1)
    from nanocaller_src import snpCaller
    snpCaller.get_snp_testing_candidates = _count_wrapper(snpCaller.get_snp_testing_candidates, counter)
    snpCaller.caller(params, q, cq, files)


def _ref_indel_worker(params, indel_dict, q, cq, files, counter):
    import torch
    torch.set_num_threads(1)
    from nanocaller_src import indelCaller
    indelCaller.get_indel_testing_candidates = _count_wrapper(indelCaller.get_indel_testing_candidates, counter)
    indelCaller.indel_run(params, indel_dict, q, cq, files)


class ReferenceWorkers:
    """The reference's own multiprocessing CPU path: `cores` processes running the UNMODIFIED `snpCaller.caller`
    (snpCaller.py:57-198) — and for mode=all `indelCaller.indel_run` (indelCaller.py:41-189) — started the way `call_manager`
    starts them (snpCaller.py:213-245, indelCaller.py:286-350), chunks pulled from a shared queue.  The modules are the
    byte-compiled reference (oracle/build_ref.py -> oracle/_ref); htslib / TensorFlow / MUSCLE / parasail, which exist
    neither here nor on the GPU box, are the stand-ins of oracle/shim."""

    def __init__(self, rs, cores):
        from oracle import build_ref
        build_ref.activate()
        import pysam                                 # oracle/shim/pysam.py
        pysam.unregister_all()
        pysam.register("mem://bam", rs)
        import multiprocessing as mp
        self.mp = mp.get_context("fork")
        self.cores = cores
        self.mgr = self.mp.Manager()

    def run(self, snp_chunks, indel_chunks=()):
        """-> (candidate sites, seconds)"""
        tmp = tempfile.mkdtemp(prefix="nc_ref_")
        counter = self.mp.Value("q", 0)
        t0 = time.perf_counter()
        if snp_chunks:
            q, cq, files = self.mgr.Queue(), self.mgr.Queue(), self.mgr.list()
            for c in snp_chunks:
                q.put(c)
            params = dict(DCT, sam_path="mem://bam", fasta_path="mem://bam", snp_model=MODEL, prefix="g", intermediate_snp_files_dir=tmp,
                          disable_coverage_normalization=False)
            ps = [self.mp.Process(target=_ref_snp_worker, args=(params, q, cq, files, counter)) for _ in range(min(self.cores, len(snp_chunks)))]
            for p in ps:
                p.start()
            for p in ps:
                p.join()
            if any(p.exitcode != 0 for p in ps):
                raise RuntimeError("reference SNP worker failed")
        if indel_chunks:
            q, cq, files, idict = self.mgr.Queue(), self.mgr.Queue(), self.mgr.list(), self.mgr.dict()
            for c in indel_chunks:
                q.put(("indel", dict(c, sam_path="mem://bam")))
            params = dict(IDCT, fasta_path="mem://bam", indel_model=MODEL, prefix="g", intermediate_indel_files_dir=tmp)
            ps = [self.mp.Process(target=_ref_indel_worker, args=(params, idict, q, cq, files, counter)) for _ in range(min(self.cores, len(indel_chunks)))]
            for p in ps:
                p.start()
            for p in ps:
                p.join()
            if any(p.exitcode != 0 for p in ps):
                raise RuntimeError("reference indel worker failed")
        dt = time.perf_counter() - t0
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
        return int(counter.value), dt


def cpu_arm(kind, rs, length, cores, budget_s):
    """One bounded sample of the workload on `cores` host cores.  -> (runner, sample description, kind string): runner() gives
    (sites, seconds).  The unmodified reference when oracle/_ref is built, else the oracle port."""
    chrom = WORKLOADS[kind]["chrom"]
    use_ref = reference_available()
    # measured in the build container: the reference's SNP worker handles ~125 sites/s/core on 100 kb chunks (190 on full 500 kb
    # chunks: the +-50 kb pileup flank is paid per chunk), its indel worker over the stand-in MUSCLE ~4 sites/s/core; the port
    # ~800 SNP sites/s/core, ~3 indel sites/s/core.  Size the per-core sub-chunk to the time budget.
    snp_bp_per_s = (125 if use_ref else 800) / 0.0097
    if kind == "snps":
        size = int(min(500_000, max(20_000, budget_s * snp_bp_per_s)))
        size = min(size, max(10_000, length // 4))
        snp_s, ind_s = sample_chunks(chrom, length, cores, size), []
        what = "%d chunks of %d bp per step (one per core) of the %d bp contig" % (cores, size, length)
    else:
        # indel chunks: ~190 s per 100 kb through the reference worker (ten pileup-string calls per column in the stand-in pysam, three
        # forks of the stand-in MUSCLE per site), ~20 s per 100 kb through the port
        isize = int(min(100_000, max(5_000, budget_s * 0.8 * (530 if use_ref else 5_000))))
        isize = min(isize, max(5_000, length // 4))
        size = int(min(500_000, max(20_000, budget_s * 0.2 * snp_bp_per_s)))
        size = min(size, max(10_000, length // 4))
        snp_s, ind_s = sample_chunks(chrom, length, cores, size), sample_chunks(chrom, length, cores, isize)
        what = "%d SNP chunks of %d bp + %d indel chunks of %d bp per step (one each per core) of the %d bp contig" % (cores, size, cores, isize, length)
    if use_ref:
        rw = ReferenceWorkers(rs, cores)
        return (lambda: rw.run(snp_s, ind_s)), what + ", unmodified reference workers (snpCaller.caller%s; oracle/_ref) over the pysam / TensorFlow%s stand-ins of oracle/shim" % (
            " + indelCaller.indel_run" if ind_s else "", " / MUSCLE / parasail" if ind_s else ""), "reference"
    pool = make_pool(rs, cores, with_indel=bool(ind_s))

    def run_port():
        t = time.perf_counter()
        sites = sum(pool.map(_cpu_chunk, [(c,) for c in snp_s], chunksize=1))
        if ind_s:
            sites += sum(pool.map(_cpu_indel_chunk, [(c,) for c in ind_s], chunksize=1))
        return sites, time.perf_counter() - t
    return run_port, what + ", numpy/torch-CPU restatement (oracle/)", "port"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path on every host core (see ReferenceWorkers), one process per core pulling
    chunks like snpCaller.call_manager (snpCaller.py:213-245); when oracle/_ref was not built, the oracle port."""
    if rank != 0:
        return
    kind = args.workload
    length = workload_len(kind)
    cores = os.cpu_count() or 1
    rs, _, _ = make_workload(kind, 0, length)
    total = max(1, args.steps + args.warmup)
    runner, sample, how = cpu_arm(kind, rs, length, cores, 200.0 / total)
    for _ in range(args.warmup):
        runner()
    sites = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s, _ = runner()
        sites += s
    dt = time.perf_counter() - t0
    v = sites / dt if dt > 0 else 0.0
    print(json.dumps({
        "impl": "reference", "metric": "candidate sites/sec (pileup+CNN)", "value": v, "unit": "sites/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(kind, length), "sample": sample},
        "cpu_baseline": {"value": v, "unit": "sites/s", "cores": cores, "kind": how, "sample": sample},
        "e2e": {"value": v, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


# ------------------------------------------------------------------------------------------------ in-run parity (untimed)
def parity_snp(ctx, rs, chunks, pool, cores, n_check=2):
    """Chunks of the LAST timed step against the oracle: positions, depths, strand counts, the int16 tensors bit for bit, the
    chunk mean depth, and the probabilities of the timed forward within 1e-4."""
    from nanocaller_b200.host import capi
    from oracle import cnn_oracle, snp_oracle
    n = ctx.n_sites
    meta = np.empty(n, capi.META_DTYPE)
    probs = np.empty((n, 4), np.float32)
    ctx.fetch_calls(probs, meta)
    _, _, depth, count = ctx.snp_fetch(want_mat=False)
    first = np.concatenate([[0], np.cumsum(count)])
    pick = sorted({min(len(chunks) - 1, 1), len(chunks) // 2})[:n_check]
    out = {"chunks": [], "sites": 0, "tensors_bit_exact": True, "meta_exact": True, "depth_exact": True, "max_abs_dp": 0.0}
    for ci in pick:
        c = chunks[ci]
        edges = np.linspace(c["start"], c["end"] + 1, cores + 1).astype(int)
        subs = [dict(c, start=int(a), end=int(b) - 1) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        parts = [p for p in pool.map(_parity_snp_sub, [(s,) for s in subs], chunksize=1) if p is not None]
        if not parts:
            continue
        pos = np.concatenate([p[0] for p in parts]); ref = np.concatenate([p[1] for p in parts]); mat = np.concatenate([p[2] for p in parts])
        dp = np.concatenate([p[3] for p in parts]); fwd = np.concatenate([p[7] for p in parts]); rev = np.concatenate([p[8] for p in parts])
        wdepth = round(sum(p[5] for p in parts)) / sum(p[6] for p in parts)
        lo, cnt = int(first[ci]), int(count[ci])
        m = meta[lo:lo + cnt]
        ok_meta = cnt == len(pos) and np.array_equal(m["pos"], pos) and np.array_equal(m["dp"], dp) and np.array_equal(m["fwd"], fwd) and \
            np.array_equal(m["rev"], rev) and np.array_equal(m["ref_code"], np.argmax(ref, 1))
        out["meta_exact"] &= bool(ok_meta)
        if ok_meta:
            got = ctx.snp_fetch_range(lo, cnt)[:, :1025].reshape(cnt, 5, 41, 5)
            out["tensors_bit_exact"] &= bool(np.array_equal(got, mat))
            out["depth_exact"] &= bool(float(depth[ci]) == float(wdepth))
            x = snp_oracle.scale_counts(mat.astype(np.int32), _G["tc"], coverage=float(wdepth))
            want = np.concatenate([cnn_oracle.snp_probs(_G["tensors"], x[b:b + 2000], ref[b:b + 2000].astype(np.float32)) for b in range(0, cnt, 2000)])
            out["max_abs_dp"] = max(out["max_abs_dp"], float(np.abs(want - probs[lo:lo + cnt]).max()))
        out["chunks"].append(int(ci)); out["sites"] += int(cnt)
    out["ok"] = bool(out["sites"] > 0 and out["tensors_bit_exact"] and out["meta_exact"] and out["depth_exact"] and out["max_abs_dp"] < 1e-4)
    return out


def parity_indel(ctx, ichunks, pool, imeta, iprobs, n_check=2):
    """Indel chunks of the LAST timed step against the oracle: key positions, the three float tensors bit for bit, probabilities <= 1e-4."""
    from nanocaller_b200.host import indel_pileups
    kept = indel_pileups.kept_sites(imeta, False)
    with_sites = np.unique(imeta["chunk"][kept])
    if len(with_sites) == 0:
        return {"ok": False, "sites": 0}
    pick = sorted({int(with_sites[len(with_sites) // 3]), int(with_sites[(2 * len(with_sites)) // 3])})[:n_check]
    res = pool.map(_parity_indel_chunk, [(ichunks[ci],) for ci in pick], chunksize=1)
    out = {"chunks": pick, "sites": 0, "positions_exact": True, "tensors_bit_exact": True, "max_abs_dp": 0.0}
    for ci, (pos, x, p, alleles, phase) in zip(pick, res):
        sel = np.nonzero((imeta["chunk"] == ci) & kept)[0]
        same = list(imeta["pos"][sel]) == list(pos)
        out["positions_exact"] &= bool(same)
        if same and len(sel):
            lo, hi = int(sel[0]), int(sel[-1]) + 1
            t = ctx.indel_fetch_range(lo, hi - lo)[sel - lo].reshape(len(sel), 15, 128, 2)
            out["tensors_bit_exact"] &= bool(np.array_equal(t, x))
            out["max_abs_dp"] = max(out["max_abs_dp"], float(np.abs(iprobs[sel] - p).max()))
        out["sites"] += len(sel)
    out["ok"] = bool(out["sites"] > 0 and out["positions_exact"] and out["tensors_bit_exact"] and out["max_abs_dp"] < 1e-4)
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def measure(kind, args, rank, world, local, dist, with_cpu=True, pipelined_ok=True):
    """One workload on this rank's GPU -> (result dict on rank 0 | None)."""
    import torch
    from nanocaller_b200.host import capi, indel_pileups, weights as W
    from nanocaller_b200.host.gather import gather_records, _DeviceBytes, RECORD_BYTES
    length = workload_len(kind)
    rs, chunks, ichunks = make_workload(kind, rank, length)
    ch = [(c["start"], c["end"]) for c in chunks]
    ich = [(c["start"], c["end"]) for c in ichunks]
    do_indel = kind == "all"
    ctx = capi.Context(local)
    tensors, meta = W.load_model("snp", MODEL)
    snp_blob = W.pack_snp_blob(tensors, False)
    ctx.load_snp_weights(snp_blob, meta["train_coverage"], False)
    indel_blob = None
    if do_indel:
        indel_blob = W.pack_indel_blob(W.load_model("indel", MODEL)[0])
        ctx.load_indel_weights(indel_blob, False)
    params = capi.snp_params(DCT, "diploid")
    iparams = capi.indel_params(IDCT, False) if do_indel else None

    # host buffers in pinned memory (the BAM-native arrays a reader would hand over)
    keep, arrs = [], []
    for a in (rs.pos, rs.flag, rs.cigar_off, rs.cigar, rs.seq_off, rs.l_seq, rs.seq4, rs.ref):
        t, v = pinned_copy(a)
        keep.append(t); arrs.append(v)
    tags = []
    if do_indel:
        for a in (rs.hp, rs.ps):
            t, v = pinned_copy(a)
            keep.append(t); tags.append(v)
    h2d = int(sum(a.nbytes for a in arrs) + sum(a.nbytes for a in tags))

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.sync(); torch.cuda.synchronize()

    def kernels(c, fetch):
        """One pass of the hot path on context c.  -> (snp sites, indel sites, (imeta, cns) or None)"""
        n = c.snp_scan(params, ch)
        c.snp_forward(normalize=True, impl=args.cnn_impl, fetch=False)
        if not do_indel:
            return n, 0, None
        variants = c.indel_scan(iparams, ich)
        sites = indel_pileups.order_variants(variants)               # dict semantics of `variants` (generate_indel_pileups.py:268-320), host
        got = c.indel_build(iparams, ich, sites, want_tensors=False, fetch=fetch)
        c.indel_forward(impl=args.indel_impl, fetch=False)
        return n, len(sites), got

    def step_resident():
        ctx.invalidate_decode()
        return kernels(ctx, False)

    # ---- device-resident figure
    ctx.stage_arrays(*arrs)
    if do_indel:
        ctx.stage_tags(*tags)
    for _ in range(args.warmup):
        n_sites, n_isites, _ = step_resident()
    l0 = ctx.timings()["launches"]
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ctx.event_record(0)
    acc = {"decode_ms": 0.0, "scan_ms": 0.0, "tensor_ms": 0.0, "cnn_ms": 0.0, "cnn_a_ms": 0.0}
    iacc = {"scan_ms": 0.0, "reads_ms": 0.0, "align_ms": 0.0, "msa_ms": 0.0, "allele_ms": 0.0, "cnn_ms": 0.0}
    itm = None
    for _ in range(args.steps):
        n_sites, n_isites, _ = step_resident()
        tm = ctx.timings()
        for k in acc:
            acc[k] += tm[k]
        if do_indel:
            itm = ctx.indel_timings()
            for k in iacc:
                iacc[k] += itm[k]
    ctx.event_record(1)
    dev_ms = ctx.event_elapsed_ms(0, 1)
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = ctx.timings()["launches"] - l0
    tbytes = ctx.timings()["tensor_bytes"]
    scan_bytes = ctx.timings()["scan_bytes"]
    # SURVEY 8(d) algorithmic bytes of the tensor build: 5 B per (sampled read, real column) + 1025 int16 + 64 B of site data,
    # from the site records of the last step (untimed); the candidates' (read, site) pairs complete B1 of the scan
    survey_bytes = None
    res_meta = res_probs = None
    try:
        res_meta = np.empty(max(1, int(n_sites)), capi.META_DTYPE)
        res_probs = np.empty((max(1, int(n_sites)), 4), np.float32)
        ctx.fetch_calls(res_probs, res_meta)
        res_meta, res_probs = res_meta[:int(n_sites)], res_probs[:int(n_sites)]
        _cols = res_meta["n_left"].astype(np.int64) + res_meta["n_right"].astype(np.int64) + 1
        survey_bytes = int((5 * res_meta["sample_depth"].astype(np.int64) * _cols).sum() + int(n_sites) * (1025 * 2 + 64))
        scan_bytes += int(5 * res_meta["dp"].astype(np.int64).sum())
    except Exception as _e:                                   # a reporting extra must never cost the bench line
        sys.stderr.write("bench: survey byte count skipped: %r\n" % (_e,))
    # what the timed step produced: a checksum of the call records, and the oracle on sampled chunks (below, untimed)
    checksum = None
    if res_meta is not None:
        import zlib
        checksum = "%08x" % (zlib.crc32(np.round(res_probs.astype(np.float64), 3).tobytes(), zlib.crc32(res_meta["pos"].tobytes())) & 0xffffffff)
    imeta = iprobs = None
    if do_indel:
        imeta, _, _ = ctx.indel_fetch(want_tensors=False)
        iprobs = np.empty((len(imeta), 4), np.float32)
        ctx.indel_fetch_probs(iprobs)

    # ---- end-to-end figure, serial: stage -> kernels -> fetch, one contig after the other
    pin = {}

    def pinned(name, shape, dtype):
        need = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if name not in pin or pin[name].nbytes < need:
            t = torch.empty(int(need * 1.2) + 64, dtype=torch.uint8, pin_memory=True)
            keep.append(t); pin[name] = t.numpy()
        return pin[name][:need].view(dtype).reshape(shape)

    rec_parts = []          # N > 1: every step's call records stay on the device until the run's one gather

    def fetch_results(c, n, ni, tag, got):
        """D2H of the step's results (N = 1) or hand-over to the run's gather (N > 1).  -> bytes copied to the host"""
        b = 0
        if world > 1:
            c.sync()
            _, meta_ptr, probs_ptr, nn = c.device_buffers()
            if nn:
                pr = torch.as_tensor(_DeviceBytes(probs_ptr, nn * 16), device="cuda").view(nn, 16)
                me = torch.as_tensor(_DeviceBytes(meta_ptr, nn * 40), device="cuda").view(nn, 40)
                rec_parts.append(torch.cat([pr, me], 1))
                torch.cuda.current_stream().synchronize()        # the library reuses these buffers in the next step
        elif n:
            c.fetch_calls(pinned(tag + "p", (n, 4), np.float32), pinned(tag + "m", (n, capi.META_DTYPE.itemsize), np.uint8))
            b += n * (16 + capi.META_DTYPE.itemsize)
        if do_indel and ni:
            c.indel_fetch_probs(pinned(tag + "ip", (ni, 4), np.float32))
            b += ni * 16 + got[0].nbytes + got[2].nbytes
        return b

    allele_s = [0.0]

    def step_e2e(c=None, tag="a"):
        c = c or ctx
        c.stage_arrays(*arrs)
        if do_indel:
            c.stage_tags(*tags)
        n, ni, got = kernels(c, True)
        b = fetch_results(c, n, ni, tag, got)
        if do_indel and ni:
            t = time.perf_counter()
            indel_pileups.AllelePredictions(rs, IDCT, got[0], got[2], False, device_lengths=c.indel_fetch_alleles())   # I4 ran on the GPU after msa; this gathers the windows
            allele_s[0] += time.perf_counter() - t
        return n + ni, b

    def gather_run():
        """The run's one exchange: all steps' call records to rank 0 (pinned host memory)."""
        if world == 1:
            return 0
        rec = torch.cat(rec_parts, 0) if rec_parts else torch.zeros((0, RECORD_BYTES), dtype=torch.uint8, device="cuda")
        rec_parts.clear()
        host, counts = gather_records(rec, dist, world, rank=rank, to_host=True)
        return int(sum(counts)) * RECORD_BYTES

    for _ in range(args.warmup):
        step_e2e()
    gather_run()
    barrier()
    allele_s[0] = 0.0
    t0 = time.perf_counter()
    d2h_total = 0
    for _ in range(args.steps):
        n_e2e, b = step_e2e()
        d2h_total += b
    ctx.sync()
    d2h_total += gather_run()
    e2e_serial_s = time.perf_counter() - t0
    allele_ms = allele_s[0] / args.steps * 1e3
    barrier()

    # ---- end-to-end figure, pipelined: two contexts (two streams) per GPU, each looping over whole steps, so the H2D copy of one
    #      contig overlaps the kernels of the other — how a run over many contigs is driven.  Same work per step, same API calls.
    #      With N > 1 the steps' call records stay on the device and the run ends with ONE gather to rank 0 (the reference
    #      concatenates its workers' files once, snpCaller.py:258-280), inside the timed region.
    e2e_s = e2e_serial_s
    pipelined = False
    dev_gb = (h2d * 2.6 + n_sites * (2064 + 10240 + 6912 + 100)) / 1e9         # rough footprint of one context
    if args.steps >= 2 and pipelined_ok and 2 * dev_gb < 120:
        import threading
        ctx2 = capi.Context(local)
        ctx2.load_snp_weights(snp_blob, meta["train_coverage"], False)
        if do_indel:
            ctx2.load_indel_weights(indel_blob, False)
        ctxs = [ctx, ctx2]
        blocking = world >= 4                           # 2 x N host threads on one host: sleeping waits beat spinning ones from 4 ranks on
        for c_ in ctxs:
            c_.set_blocking_sync(blocking)
        aff = os.sched_getaffinity(0)

        def worker(i, steps_of_i, out):
            torch.cuda.set_device(local)                   # the current device is per host thread
            os.sched_setaffinity(0, aff)
            for _ in steps_of_i:
                s, b = step_e2e(ctxs[i], "ab"[i])
                out[i] += s
                out[2 + i] += b

        def run(nsteps_total):
            out = [0, 0, 0, 0]
            th = [threading.Thread(target=worker, args=(i, list(range(i, nsteps_total, 2)), out)) for i in range(2)]
            barrier()
            t = time.perf_counter()
            for x in th:
                x.start()
            for x in th:
                x.join()
            for c in ctxs:
                c.sync()
            b = gather_run()
            dt = time.perf_counter() - t
            barrier()
            return out[0] + out[1], dt, out[2] + out[3] + b

        run(max(2, args.warmup))
        sites_p, e2e_p, d2h_p = run(args.steps)
        if sites_p == (n_sites + n_isites) * args.steps:
            e2e_s, pipelined, d2h_total = e2e_p, True, d2h_p
        ctx2.close()
        ctx.set_blocking_sync(False)

    tot_sites, dev_ms_max, e2e_max = n_sites + n_isites, dev_ms, e2e_s
    if world > 1:
        t = torch.tensor([float(n_sites + n_isites), dev_ms, e2e_s, e2e_serial_s, float(d2h_total)], dtype=torch.float64, device="cuda")
        ts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        tot_sites = int(sum(x[0].item() for x in ts)); dev_ms_max = max(x[1].item() for x in ts); e2e_max = max(x[2].item() for x in ts)
        e2e_serial_s = max(x[3].item() for x in ts)
        d2h_total = int(ts[0][4].item())                # rank 0 receives the gathered records
    if rank != 0:
        ctx.close()
        return None

    hbm, tflops, which = peaks()
    ms_per_step = dev_ms_max / args.steps
    value = tot_sites / (ms_per_step * 1e-3)
    ph = {k: v / args.steps for k, v in acc.items()}
    cnn_ms, tensor_ms, ta_ms = ph["cnn_ms"], ph["tensor_ms"], ph["cnn_a_ms"]
    ach = FLOP_PER_SITE * n_sites / (cnn_ms * 1e-3) / 1e12
    tr = measured_traffic() or {}
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    roof = {"kernel": "CNN forward (tc_trunk_a + tc_trunk_b + tc_fc: tcgen05 kind::f16, fp16 hi/lo split operands, hi*hi + lo*hi + hi*lo)",
            "bound": "tensor", "achieved": ach, "peak": tflops, "unit": "TFLOP/s", "frac": ach / tflops,
            "traffic": (tr["cnn_dram_bytes_per_site"] * n_sites) if "cnn_dram_bytes_per_site" in tr else None,
            "traffic_source": tr.get("source") or tr.get("file"),
            "algorithmic_bytes": n_sites * CNN_BYTES_PER_SITE,
            "algorithmic_bytes_note": "SURVEY 8(d): 2,050 B int16 tensor in + 16 B out per site; activations that cross HBM between the kernels are traffic, not algorithmic bytes",
            "peak_source": which + ", sustained bf16", "flop_per_site": FLOP_PER_SITE, "ms_per_launch_group": cnn_ms}
    if ta_ms > 0:
        ta_flop = 2 * (82_000 + 82_000 + 410_000 + 737_280)                 # conv1_1 + conv1_2 + conv1_3 + conv2 MACs per site
        ta_ach = ta_flop * n_sites / (ta_ms * 1e-3) / 1e12
        roof["dominant_kernel"] = {"kernel": "tc_trunk_a_kernel (conv1_1 + conv1_2 + conv1_3 + conv2 of every site; %.0f %% of the step)" % (100 * ta_ms / ms_per_step),
                                   "ms_per_launch": ta_ms, "flop_per_site": ta_flop, "achieved": ta_ach, "peak": tflops, "unit": "TFLOP/s",
                                   "frac": ta_ach / tflops,
                                   "traffic": (tr.get("per_kernel", {}).get("tc_trunk_a_kernel", {}).get("dram_bytes", 0) / max(1, tr.get("sites", 1)) * n_sites) or None,
                                   "algorithmic_bytes": n_sites * 2050}
    roof2 = {"kernel": "tensor_kernel (K2 pileup tensor build)", "bound": "hbm", "achieved": tbytes / (tensor_ms * 1e-3) / 1e9,
             "peak": hbm, "unit": "GB/s", "frac": tbytes / (tensor_ms * 1e-3) / 1e9 / hbm,
             "traffic": (tr["tensor_dram_bytes_per_site"] * n_sites) if "tensor_dram_bytes_per_site" in tr else None,
             "bytes_per_site": tbytes / max(1, n_sites), "ms_per_launch": tensor_ms,
             "note": "achieved counts the bytes this kernel has to move (int16 tensor + site record); survey_8d counts SURVEY 8(d)'s "
                     "B2 = 5 B per (sampled read, real column) + 1025 int16 + 64 B, i.e. the read-code lists a per-site kernel would stream"}
    if survey_bytes:
        roof2["survey_8d"] = {"bytes_per_site": survey_bytes / max(1, n_sites), "achieved": survey_bytes / (tensor_ms * 1e-3) / 1e9,
                              "frac": survey_bytes / (tensor_ms * 1e-3) / 1e9 / hbm}
    ds_ms = ph["decode_ms"] + ph["scan_ms"]
    roof3 = {"kernel": "K0 decode (cigar_scan, seq_codes, row_fill) + K1 scan (scan_kernel, site lists, neighbour matrix)", "bound": "hbm",
             "achieved": scan_bytes / (ds_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": scan_bytes / (ds_ms * 1e-3) / 1e9 / hbm,
             "algorithmic_bytes": int(scan_bytes), "ms_per_launch_group": ds_ms,
             "traffic": (tr["scan_dram_bytes_per_base"] * rs.aligned_bases()) if "scan_dram_bytes_per_base" in tr else None,
             "note": "SURVEY 8(d) B1: 4 B per CIGAR op + 4-bit bases + 16 B per read + 1 B per piled reference position + 5 B per (read, kept site)"}

    out = {"metric": "candidate sites/sec (pileup+CNN)", "value": value, "unit": "sites/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "int16 pileup counts; CNN fp16 hi/lo split operands with fp32 accumulation" if args.cnn_impl == 0 else "int16 pileup counts; f32 CNN",
           "data": "synthetic",
           "config": {"workload": workload_name(kind, length), "snp_sites_per_gpu": int(n_sites), "indel_sites_per_gpu": int(n_isites),
                      "chunks": len(ch), "indel_chunks": len(ich), "aligned_bases_per_gpu": rs.aligned_bases(), "reads_per_gpu": rs.n,
                      "l2": "inputs (%.2f GB) and tensors (%.2f GB) exceed the 126 MB L2, no flush needed" % (h2d / 1e9, n_sites * 2064 / 1e9),
                      "parallelism": "1 rank per GPU, one contig per rank, one gather of the run's call records to rank 0" if world > 1 else "single GPU"},
           "phase_ms": ph,
           "e2e": {"value": tot_sites / (e2e_max / args.steps), "unit": "sites/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h_total / args.steps),
                   "ms_per_step": e2e_max / args.steps * 1e3,
                   "mode": ("2 contexts per GPU, H2D of one contig overlapped with kernels of the other"
                            + ("; the run ends with one NCCL gather of all steps' call records to rank 0, inside the timed region" if world > 1 else "")) if pipelined else "serial",
                   "host_waits": "blocking (cudaEventBlockingSync)" if (pipelined and world >= 4) else "spinning",
                   "serial_value": tot_sites / (e2e_serial_s / args.steps)},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_pileup": roof2, "roofline_scan": roof3,
           "result_checksum": checksum}
    if do_indel:
        iph = {k: v / args.steps for k, v in iacc.items()}
        out["indel_phase_ms"] = iph
        out["e2e"]["allele_gather_host_ms"] = allele_ms
        n_ent = int(itm["n_entries"]) if itm else 0
        wa = 161
        cells = n_ent * wa * wa                                 # full DP table per (slice, reference window) pair
        out["roofline_indel"] = {
            "kernel": "indel build (indel_site_reads + indel_align + indel_msa + indel_allele)", "bound": "hbm", "peak": hbm, "unit": "GB/s",
            "algorithmic_bytes": int(itm["build_bytes"]) if itm else None,
            "ms_per_launch_group": iph["reads_ms"] + iph["align_ms"] + iph["msa_ms"] + iph["allele_ms"],
            "achieved": (itm["build_bytes"] / ((iph["reads_ms"] + iph["align_ms"] + iph["msa_ms"] + iph["allele_ms"]) * 1e-3) / 1e9) if itm else None,
            "traffic": (tr["indel_build_dram_bytes_per_site"] * n_isites) if "indel_build_dram_bytes_per_site" in tr else None,
            "per_kernel_ms": {"indel_scan (depth, events, windows, decide, greedy)": iph["scan_ms"], "indel_site_reads": iph["reads_ms"],
                              "indel_align": iph["align_ms"], "indel_msa": iph["msa_ms"], "indel_allele (consensus x reference NW + allele walk)": iph["allele_ms"],
                              "indel CNN": iph["cnn_ms"]},
            "aligned_slices": n_ent, "dp_cell_updates_per_s": cells / (iph["align_ms"] * 1e-3) if iph["align_ms"] > 0 else None,
            "note": "the alignment is integer dynamic programming (161 x 161 cells per slice): neither HBM nor the tensor pipe bounds it; "
                    "bytes = slices + reference windows read, tensors + consensus written"}
        out["roofline_indel"]["frac"] = (out["roofline_indel"]["achieved"] / hbm) if out["roofline_indel"]["achieved"] else None
        icnn = FLOP_PER_INDEL_SITE * n_isites / (iph["cnn_ms"] * 1e-3) / 1e12 if iph["cnn_ms"] > 0 else 0.0
        out["roofline_indel_cnn"] = {"kernel": "Indel_model forward (%s)" % ("tcgen05" if (args.indel_impl == 0 and INDEL_CNN_ON_TENSOR_CORES) else "conv_f32_kernel, fp32 CUDA cores"), "bound": "tensor",
                                     "achieved": icnn, "peak": tflops, "unit": "TFLOP/s", "frac": icnn / tflops, "flop_per_site": FLOP_PER_INDEL_SITE,
                                     "ms_per_launch_group": iph["cnn_ms"], "algorithmic_bytes": n_isites * (15 * 128 * 2 * 4 + 16)}
        isc = int(itm["scan_bytes"]) if itm else 0
        out["roofline_indel_scan"] = {"kernel": "nc_indel_scan", "bound": "hbm", "achieved": isc / (iph["scan_ms"] * 1e-3) / 1e9 if iph["scan_ms"] > 0 else None,
                                      "peak": hbm, "unit": "GB/s", "algorithmic_bytes": isc, "ms_per_launch_group": iph["scan_ms"]}
        if out["roofline_indel_scan"]["achieved"]:
            out["roofline_indel_scan"]["frac"] = out["roofline_indel_scan"]["achieved"] / hbm

    # ---- untimed: what the timed step produced against the oracle on sampled chunks, and the CPU baseline
    if with_cpu and world == 1 and os.environ.get("NC_BENCH_NO_CPU") != "1":       # NC_BENCH_NO_CPU=1: profiler runs
        cores = os.cpu_count() or 1
        try:
            pool = make_pool(rs, cores, with_indel=do_indel)
            ctx.invalidate_decode()
            kernels(ctx, False)                                  # the state of a resident step (the e2e loop re-staged the same arrays)
            par = {"snp": parity_snp(ctx, rs, chunks, pool, cores)}
            if do_indel:
                par["indel"] = parity_indel(ctx, ichunks, pool, imeta, iprobs)
            pool.close()
            par["ok"] = all(v.get("ok") for v in par.values())
            out["parity"] = par
        except Exception as e:
            out["parity"] = {"ok": False, "error": repr(e)}
        try:
            runner, sample, how = cpu_arm(kind, rs, length, cores, 20.0)
            s, dt = runner()
            out["cpu_baseline"] = {"value": s / dt, "unit": "sites/s", "cores": cores, "kind": how, "sample": sample + ", %.1f s" % dt}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "unit": "sites/s", "cores": cores, "kind": "port", "sample": "failed: %r" % (e,)}
    else:
        out["cpu_baseline"] = None
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="snps", choices=["snps", "all"], help="snps: BASELINE configs[1] (default); all: configs[2]")
    ap.add_argument("--no-all", action="store_true", help="N = 1 default run: skip the nested configs[2] measurement")
    ap.add_argument("--cnn-impl", type=int, default=0, help="SNP CNN: 0 tcgen05 (default), 1 fp32 CUDA cores")
    ap.add_argument("--indel-impl", type=int, default=0, help="indel CNN: 0 tcgen05 where built (default), 1 fp32 CUDA cores")
    ap.add_argument("--from-bam", type=int, default=20_000_000, metavar="BP",
                    help="also report sites/s from a BAM + FASTA on disk (SURVEY 8d figure ii) for a synthetic contig of BP bases (0: skip): the device "
                         "reader (compressed bytes over PCIe, BGZF inflate + record decoding on the GPU) and, beside it, the host reader (zlib)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.warmup = max(args.warmup, 3)

    out = measure(args.workload, args, rank, world, local, dist)
    if rank == 0 and numa:
        out["config"]["host_binding"] = numa
    if rank == 0 and args.workload == "snps" and world == 1 and not args.no_all and os.environ.get("NC_BENCH_NO_ALL") != "1":
        # configs[2] on the same GPU, nested under "mode_all": fewer steps (a step is a 250 Mb contig), serial end-to-end figure
        try:
            sub = argparse.Namespace(**vars(args))
            sub.steps = max(2, min(args.steps, 5)); sub.warmup = 3
            r = measure("all", sub, rank, world, local, dist, pipelined_ok=False)
            out["mode_all"] = {k: r[k] for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "config", "phase_ms", "indel_phase_ms", "e2e",
                                                 "gpu_launches", "roofline", "roofline_pileup", "roofline_scan", "roofline_indel", "roofline_indel_cnn",
                                                 "roofline_indel_scan", "parity", "cpu_baseline", "result_checksum") if k in r}
        except Exception as e:
            out["mode_all"] = {"error": repr(e)}
    if rank == 0 and args.from_bam > 0 and world == 1 and os.environ.get("NC_BENCH_NO_CPU") != "1":
        try:
            out["from_bam"] = from_bam_figure(args, local)
        except Exception as e:
            out["from_bam"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def from_bam_figure(args, local):
    """SURVEY 8(d) figure (ii): BAM + FASTA on disk (page cache) -> libnc_bamio -> H2D -> kernels -> D2H."""
    import shutil
    import torch
    from nanocaller_b200.cli import get_chunks
    from nanocaller_b200.host import bamio, capi, weights as W
    from nanocaller_b200.synth import make_world
    rs_d = make_world(chrom="chr20", preset="ont", contig_len=args.from_bam, seed=20, coverage=30.0).reads
    tmpd = tempfile.mkdtemp(prefix="nc_bench_")
    bam_p, fa_p = os.path.join(tmpd, "d.bam"), os.path.join(tmpd, "d.fa")
    bamio.write_bam(bam_p, [rs_d], index=True); bamio.write_fasta(fa_p, [rs_d])
    ch_d = [(c["start"], c["end"]) for c in get_chunks([("chr20", 1, args.from_bam, "diploid")], 1)]
    ctx = capi.Context(local)
    tensors, meta = W.load_model("snp", MODEL)
    ctx.load_snp_weights(W.pack_snp_blob(tensors, False), meta["train_coverage"], False)
    params = capi.snp_params(DCT, "diploid")
    keep = []
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

    bufs = {}

    def finish(n):
        if "p" not in bufs or len(bufs["p"]) < n:
            tp = torch.empty((int(n * 1.2) + 16, 4), dtype=torch.float32, pin_memory=True)
            tm = torch.empty((int(n * 1.2) + 16, capi.META_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
            keep.extend([tp, tm]); bufs["p"], bufs["m"] = tp.numpy(), tm.numpy()
        ctx.fetch_calls(bufs["p"][:n], bufs["m"][:n])
        return n

    def host_step():
        """libnc_bamio: zlib inflate + record copy on the host threads into pinned memory, then H2D of the decoded arrays"""
        arena["off"] = 0
        fasta = bamio.read_fasta(fa_p)
        sets, _ = bamio.read_bam_native(bam_p, fasta, alloc=pinned_alloc)
        r = sets[0]
        ctx.stage_arrays(r.pos, r.flag, r.cigar_off, r.cigar, r.seq_off, r.l_seq, r.seq4, r.ref)
        n = ctx.snp_scan(params, ch_d)
        ctx.snp_forward(normalize=True, impl=args.cnn_impl, fetch=False)
        return finish(n)

    tms = []

    def device_step():
        """nc_bam_device_open: the compressed bytes cross PCIe, BGZF inflate + record decoding on the GPU"""
        fasta = bamio.read_fasta(fa_p)
        ctx.bam_device_open(bam_p)
        tms.append(ctx.bam_device_timings())
        ctx.bam_device_stage(0, fasta["chr20"])
        n = ctx.snp_scan(params, ch_d)
        ctx.snp_forward(normalize=True, impl=args.cnn_impl, fetch=False)
        return finish(n)

    def timed(fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            n = fn()
        return n, (time.perf_counter() - t0) / reps

    n_h, dt_h = timed(host_step)
    n_d, dt_d = timed(device_step)
    ctx.bam_device_close()
    ctx.close()
    tm = tms[-1]
    res = {"value": n_d / dt_d, "unit": "sites/s", "ms_per_step": dt_d * 1e3, "contig_bp": args.from_bam, "sites": int(n_d),
           "bam_bytes": os.path.getsize(bam_p), "host_threads": os.cpu_count(),
           "path": "nc_bam_device_open (mmap -> pinned -> H2D of the COMPRESSED file, bgzf_inflate_kernel, record walk + decoding kernels) -> nc_bam_device_stage -> kernels -> D2H; file in the page cache",
           "device_reader_ms": {k: tm[k] for k in ("host_ms", "h2d_ms", "inflate_ms", "records_ms")}, "record_walk": tm["record_walk"],
           "inflate_GBps_out": tm["inflated_bytes"] / (tm["inflate_ms"] * 1e-3) / 1e9 if tm["inflate_ms"] > 0 else None,
           "inflated_bytes": tm["inflated_bytes"],
           "host_reader": {"value": n_h / dt_h, "unit": "sites/s", "ms_per_step": dt_h * 1e3, "sites": int(n_h),
                           "path": "libnc_bamio (parallel zlib inflate + record copy into pinned memory) -> nc_stage_reads -> kernels -> D2H"},
           "same_sites": bool(n_h == n_d)}
    shutil.rmtree(tmpd, ignore_errors=True)
    return res


if __name__ == "__main__":
    main()
