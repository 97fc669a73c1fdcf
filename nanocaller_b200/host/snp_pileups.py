"""Host mirror of nanocaller_src/generate_SNP_pileups.py over the CUDA library.

`get_snp_testing_candidates(dct, region)` keeps the reference signature and return tuple
(generate_SNP_pileups.py:103, :278; call site snpCaller.py:86) so the worker code above it reads like
the reference's.  All work happens in libnanocaller_b200.so (K0 decode, K1 scan, K2 tensor build);
this module only resolves paths, stages arrays and reshapes the results."""
import numpy as np

from . import capi, sources

_ctx_cache = {}
_staged = {}


def context(device=0):
    """Per-device library context (one per process, like one reference worker)."""
    if device not in _ctx_cache:
        _ctx_cache[device] = capi.Context(device)
    return _ctx_cache[device]


def reset():
    for c in _ctx_cache.values():
        c.close()
    _ctx_cache.clear()
    _staged.clear()


def stage(ctx, rs):
    """Upload a contig unless it is the one already staged on this context."""
    key = id(ctx)
    if _staged.get(key) is not rs:
        if isinstance(rs, sources.DeviceContig):          # reads already on the device (nc_bam_device_open): no host copy exists
            ctx.bam_device_stage(rs.index, rs.ref)
        else:
            ctx.stage_reads(rs)
        _staged[key] = rs


def scan_chunks(ctx, rs, dct, chunks, ploidy, bed=None):
    """Run K0-K2 for a list of chunk dicts of one contig.  Results stay on the device; returns n_sites."""
    stage(ctx, rs)
    params = capi.snp_params(dct, ploidy)
    return ctx.snp_scan(params, [(c["start"], c["end"]) for c in chunks], bed)


def unpack(mat, meta, depth, count, n_chunks):
    """Split fetched device results into per-chunk reference-shaped tuples."""
    out = []
    off = 0
    for ci in range(n_chunks):
        n = int(count[ci])
        if n == 0:
            out.append(([], [], [], [], [], 0, [], []))      # generate_SNP_pileups.py:193-197,278
            continue
        m = meta[off:off + n]
        x = mat[off:off + n, :capi.SITE_ELEMS].reshape(n, 5, 41, 5).astype(np.float32)      # :266
        ref = np.zeros((n, 4), np.int32)
        ref[np.arange(n), m["ref_code"]] = 1                                                  # :269-270
        out.append((m["pos"].astype(np.int64), ref, x, m["dp"].astype(np.int64),
                    m["alt"].astype(np.float64) / m["dp"].astype(np.float64), np.float64(depth[ci]),
                    m["fwd"].astype(np.float64), m["rev"].astype(np.float64)))
        off += n
    return out


def get_snp_testing_candidates(dct, region, device=0):
    """Drop-in for generate_SNP_pileups.get_snp_testing_candidates (same dct keys, same 8-tuple)."""
    ctx = context(device)
    rs = sources.resolve(dct["sam_path"], region["chrom"])
    bed = sources.bed_intervals(dct.get("exclude_bed"), region["chrom"])
    scan_chunks(ctx, rs, dct, [region], region["ploidy"], bed)
    mat, meta, depth, count = ctx.snp_fetch()
    return unpack(mat, meta, depth, count, 1)[0]
