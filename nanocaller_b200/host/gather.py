"""Gather of per-site call records to rank 0 — the one exchange of the multi-GPU path.

The reference exchanges results as per-process VCF files concatenated by the parent
(snpCaller.py:258-280).  Here every rank owns a contiguous run of chunks; after the CNN each rank holds
fixed-width call records (4 probabilities + NcSiteMeta = 56 B/site) in HBM, and rank 0 needs all of
them, ordered by rank (= genomic order), before it writes the VCF.  Two collectives: an all-gather of the
record counts and a padded all-gather of the records (NCCL over NVLink on the GPU box; gloo in CPU tests)."""
import numpy as np

RECORD_BYTES = 16 + 40


class _DeviceBytes:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


_BUFS = {}


def _buffers(key, make):
    if key not in _BUFS:
        _BUFS[key] = make()
    return _BUFS[key]


def gather_records(records, dist, world, device=None, rank=None, to_host=False):
    """records: uint8 tensor [n, W] (any n per rank).
    Default: every rank gets (tensor [sum n, W] in rank order, counts) — an all-gather.
    With `rank` given: a gather to rank 0 only; rank 0 gets the tensor (pinned host memory when `to_host`), the other
    ranks get None.  Staging buffers are cached across calls (a per-step allocation costs more than the exchange)."""
    import torch
    dev = records.device if device is None else device
    W = records.shape[1]
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=dev)
    cnt = _buffers(("cnt", world, str(dev)), lambda: torch.zeros(world, dtype=torch.int64, device=dev))
    dist.all_gather(list(cnt.split(1)), n)         # the outputs are views of `cnt`
    counts = [int(c) for c in cnt.tolist()]
    nmax = max(max(counts), 1)
    cap = _BUFS.get(("cap", world, W, str(dev)), 0)
    if nmax > cap:
        cap = int(nmax * 1.25) + 16
        _BUFS[("cap", world, W, str(dev))] = cap
        for k in [k for k in _BUFS if k[0] in ("pad", "parts", "pin") and k[1:] == (world, W, str(dev))]:
            del _BUFS[k]
    pad = _buffers(("pad", world, W, str(dev)), lambda: torch.zeros((cap, W), dtype=torch.uint8, device=dev))
    pad[:records.shape[0]] = records
    if rank is None:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        return torch.cat([p[:c] for p, c in zip(parts, counts)], 0), counts
    if rank == 0:
        parts = _buffers(("parts", world, W, str(dev)), lambda: [torch.empty_like(pad) for _ in range(world)])
        dist.gather(pad, parts, dst=0)
        if not to_host:
            return torch.cat([p[:c] for p, c in zip(parts, counts)], 0), counts
        pin = _buffers(("pin", world, W, str(dev)), lambda: torch.empty((cap * world, W), dtype=torch.uint8, pin_memory=(dev.type == "cuda")))
        off = 0
        for p_, c in zip(parts, counts):
            pin[off:off + c].copy_(p_[:c], non_blocking=True)
            off += c
        if dev.type == "cuda":
            torch.cuda.current_stream().synchronize()
        return pin[:off], counts
    dist.gather(pad, None, dst=0)
    return None, counts


def gather_calls(ctx, dist, rank, world):
    """All ranks: contribute the call records of the last scan+forward; rank 0 receives them in rank (= genomic) order
    and copies them to pinned host memory.  Returns the number of gathered sites."""
    import torch
    ctx.sync()                                     # records are produced on the library's stream
    _, meta_ptr, probs_ptr, n = ctx.device_buffers()
    if n:
        probs = torch.as_tensor(_DeviceBytes(probs_ptr, n * 16), device="cuda").view(n, 16)
        meta = torch.as_tensor(_DeviceBytes(meta_ptr, n * 40), device="cuda").view(n, 40)
        rec = torch.cat([probs, meta], 1)
    else:
        rec = torch.zeros((0, RECORD_BYTES), dtype=torch.uint8, device="cuda")
    host, counts = gather_records(rec, dist, world, rank=rank, to_host=True)
    return int(sum(counts))


def split_records(host_records):
    """[n, 56] uint8 -> (probs float32 [n,4], meta structured array)."""
    from .capi import META_DTYPE
    a = np.ascontiguousarray(host_records)
    probs = a[:, :16].copy().view(np.float32).reshape(-1, 4)
    meta = a[:, 16:].copy().view(META_DTYPE).reshape(-1)
    return probs, meta
