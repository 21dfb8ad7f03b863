"""N > 1 host logic (window sharding + score gather) on CPU: world_size-2 gloo processes, 127.0.0.1 rendezvous."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plantcaduceus_b200.sharding import gather_rows, shard_range, score_sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 185, 256, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_score(ascii_rows: np.ndarray) -> np.ndarray:
    """Deterministic stand-in for the engine: a function of the window bytes only (so any mis-ordering shows)."""
    a = ascii_rows.astype(np.float32)
    return np.stack([a.sum(1), a[:, 0], a[:, -1], (a * np.arange(a.shape[1])).sum(1)], axis=1).astype(np.float32)


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        windows = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=(n, 64))
        full = score_sharded(_fake_score, windows)
        ok = np.array_equal(full, _fake_score(windows))
        # gather_rows with uneven shards
        s, e = shard_range(n, rank, world)
        g = gather_rows(torch.arange(s, e, dtype=torch.float32)[:, None], n)
        ok = ok and torch.equal(g[:, 0], torch.arange(n, dtype=torch.float32))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [185, 2, 1])
def test_score_gather_world_size_2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(2))
    assert results == {0: True, 1: True}
