"""world_size-2 gloo tests of the sharding + all-gather logic (the N>1 path of bench.py / dist.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dim_b200.dist import all_gather_codes, shard_batch, shard_range


def test_shard_range_partitions():
    for total in (1, 2, 7, 256, 2048, 2051):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_carries_global_index():
    b = {"x": torch.arange(10)[:, None].repeat(1, 3), "note": "keep"}
    s = shard_batch(b, 1, 3)
    assert s["x"][:, 0].tolist() == [4, 5, 6] and s["batch_index"].tolist() == [4, 5, 6] and s["note"] == "keep"
    assert s["batch_index"].dtype == torch.int32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, S, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * S, dtype=torch.int64).view(total, S)
        s, e = shard_range(total, rank, world)
        out = all_gather_codes(full[s:e].clone(), total)
        q.put((rank, torch.equal(out, full)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])          # even and ragged shards
def test_all_gather_codes_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
