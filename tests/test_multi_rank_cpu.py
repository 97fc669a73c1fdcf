"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path — contiguous chunk sharding and the
gather of variable-length call-record buffers to rank order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nanocaller_b200.host.gather import RECORD_BYTES, gather_records, split_records
from nanocaller_b200.host.shard import shard_chunks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _records(rank, n):
    rng = np.random.RandomState(100 + rank)
    from nanocaller_b200.host.capi import META_DTYPE
    probs = rng.rand(n, 4).astype(np.float32)
    meta = np.zeros(n, META_DTYPE)
    meta["pos"] = np.arange(n) + 1000 * rank
    meta["chunk"] = rank
    meta["dp"] = rng.randint(4, 60, n)
    rec = np.concatenate([probs.view(np.uint8).reshape(n, 16), meta.view(np.uint8).reshape(n, 40)], 1)
    return probs, meta, rec


def _worker(rank, world, port, counts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, _, rec = _records(rank, counts[rank])
    allrec, got_counts = gather_records(torch.from_numpy(rec), dist, world)
    q.put((rank, allrec.numpy().copy(), got_counts))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_records_world2_gloo():
    world, counts = 2, [7, 0]
    for counts in ([7, 3], [0, 5], [4, 4]):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        ps = [ctx.Process(target=_worker, args=(r, world, port, counts, q)) for r in range(world)]
        for p in ps:
            p.start()
        res = [q.get(timeout=120) for _ in range(world)]
        for p in ps:
            p.join(timeout=60)
            assert p.exitcode == 0
        want_p = np.concatenate([_records(r, counts[r])[0] for r in range(world)])
        want_m = np.concatenate([_records(r, counts[r])[1] for r in range(world)])
        for rank, allrec, got_counts in res:
            assert got_counts == counts
            assert allrec.shape == (sum(counts), RECORD_BYTES)
            probs, meta = split_records(allrec)
            np.testing.assert_array_equal(probs, want_p)
            np.testing.assert_array_equal(meta["pos"], want_m["pos"])
            np.testing.assert_array_equal(meta["chunk"], want_m["chunk"])


def _worker_root(rank, world, port, rounds, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for counts in rounds:                                   # consecutive steps reuse (and grow) the cached staging buffers
        _, _, rec = _records(rank, counts[rank])
        host, got_counts = gather_records(torch.from_numpy(rec), dist, world, rank=rank, to_host=True)
        out.append((None if host is None else host.numpy().copy(), got_counts))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_to_rank0_world2_gloo():
    """The bench / product exchange: gather to rank 0 only, into cached buffers, over several steps of changing sizes."""
    world = 2
    rounds = [[7, 3], [2, 9], [40, 0], [5, 5]]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_root, args=(r, world, port, rounds, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k, counts in enumerate(rounds):
        host0, c0 = res[0][k]
        host1, c1 = res[1][k]
        assert c0 == counts and c1 == counts and host1 is None
        probs, meta = split_records(host0)
        np.testing.assert_array_equal(probs, np.concatenate([_records(r, counts[r])[0] for r in range(world)]))
        np.testing.assert_array_equal(meta["pos"], np.concatenate([_records(r, counts[r])[1] for r in range(world)])["pos"])


def test_shard_chunks_contiguous_balanced_and_complete():
    from oracle.snp_oracle import get_chunks
    chunks = get_chunks([("chr1", 1, 7_300_000, "diploid"), ("chr2", 1, 2_100_000, "diploid")], 1)
    for world in (1, 2, 3, 8):
        parts = shard_chunks(chunks, world)
        flat = [i for p in parts for i in p]
        assert flat == list(range(len(chunks)))                       # every chunk exactly once, order kept
        sizes = [sum(chunks[i]["end"] - chunks[i]["start"] + 1 for i in p) for p in parts]
        assert max(sizes) - min(sizes) <= 2 * 500_001                # within two chunks of each other
    w = [1] * 10 + [100] + [1] * 10
    parts = shard_chunks([{"start": 1, "end": 1}] * 21, 2, w)
    assert sum(len(p) for p in parts) == 21 and all(parts)
