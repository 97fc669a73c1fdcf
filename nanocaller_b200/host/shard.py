"""Chunk-to-rank assignment for the multi-GPU path.

Chunks of utils.get_chunks are independent units (snpCaller.py:83-86); the chunk — not the rank — stays
the normalisation unit, so results do not depend on the GPU count.  Each rank receives a contiguous run of
chunks (genomic order is preserved across ranks, which makes the gathered call records globally sorted),
balanced by the chunks' weights (aligned bases when known, else span)."""


def shard_chunks(chunks, world, weights=None):
    """-> list of `world` lists of chunk indices, contiguous and in order; greedy prefix balancing."""
    n = len(chunks)
    if weights is None:
        weights = [c["end"] - c["start"] + 1 for c in chunks]
    total = float(sum(weights)) or 1.0
    out = [[] for _ in range(world)]
    acc = 0.0
    r = 0
    for i in range(n):
        # move to the next rank when this chunk's midpoint passes the rank's share boundary
        mid = acc + weights[i] / 2.0
        while r < world - 1 and mid > total * (r + 1) / world:
            r += 1
        out[r].append(i)
        acc += weights[i]
    return out
