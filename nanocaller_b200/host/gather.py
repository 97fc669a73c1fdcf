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


def gather_records(records, dist, world, device=None):
    """records: uint8 tensor [n, W] (any n per rank).  Returns (tensor [sum n, W] in rank order, counts list)."""
    import torch
    dev = records.device if device is None else device
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts) if counts else 0
    W = records.shape[1]
    pad = torch.zeros((max(nmax, 1), W), dtype=torch.uint8, device=dev)
    pad[:records.shape[0]] = records
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0), counts


def gather_calls(ctx, dist, rank, world):
    """All ranks: contribute the call records of the last scan+forward; rank 0 copies the gathered set to the host.
    Returns the number of gathered sites."""
    import torch
    ctx.sync()                                     # records are produced on the library's stream
    _, meta_ptr, probs_ptr, n = ctx.device_buffers()
    if n:
        probs = torch.as_tensor(_DeviceBytes(probs_ptr, n * 16), device="cuda").view(n, 16)
        meta = torch.as_tensor(_DeviceBytes(meta_ptr, n * 40), device="cuda").view(n, 40)
        rec = torch.cat([probs, meta], 1)
    else:
        rec = torch.zeros((0, RECORD_BYTES), dtype=torch.uint8, device="cuda")
    allrec, counts = gather_records(rec, dist, world)
    if rank == 0:
        host = allrec.cpu()
        return int(host.shape[0])
    return int(sum(counts))


def split_records(host_records):
    """[n, 56] uint8 -> (probs float32 [n,4], meta structured array)."""
    from .capi import META_DTYPE
    a = np.ascontiguousarray(host_records)
    probs = a[:, :16].copy().view(np.float32).reshape(-1, 4)
    meta = a[:, 16:].copy().view(META_DTYPE).reshape(-1)
    return probs, meta
