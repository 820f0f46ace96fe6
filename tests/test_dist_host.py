"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: the NCCL unique id travels through
torch.distributed, and the column partitions used by the solver cover the problem exactly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zquatev_b200 import dist as zd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        raw = zd.exchange_unique_id(zd._make_nccl_id)
        # the ranks also agree on who owns what
        n = 1000
        blocks = [zd.column_block(r, world, n) for r in range(world)]
        # ... and on the plans of the collective host-pointer solve (every rank must issue the same NCCL call sequence):
        # sub-block widths of a shard and the upload ranges, computed by each rank's own library call
        chunks = zd.host_pipeline_chunks(8192, world)
        ranges = zd.upload_ranges(16384, world)
        plan = torch.tensor(chunks + [0] * (4 - len(chunks)) + ranges, dtype=torch.int64)
        plans = [torch.zeros_like(plan) for _ in range(world)]
        dist.all_gather(plans, plan)
        assert all(torch.equal(plans[0], p) for p in plans), plans
        assert sum(chunks) == 8192 and ranges[0] == 0 and ranges[-1] == 16384
        t = torch.tensor([sum(raw) % 65521, blocks[rank][0], blocks[rank][1]], dtype=torch.int64)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        q.put((rank, raw, [g.tolist() for g in gathered]))
    finally:
        dist.destroy_process_group()


def test_unique_id_exchange_gloo_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[1][1] and len(res[0][1]) == 128 and any(res[0][1])
    g = res[0][2]
    assert g[0][0] == g[1][0]                       # same id checksum seen by both ranks
    assert g[0][1] == 0 and g[0][1] + g[0][2] == g[1][1] and g[1][1] + g[1][2] == 1000


@pytest.mark.parametrize("n,world", [(1, 2), (7, 8), (64, 8), (1000, 3), (16384, 8), (100, 1)])
def test_column_blocks_partition(n, world):
    cover = []
    for r in range(world):
        c0, nc = zd.column_block(r, world, n)
        assert 0 <= c0 <= n and nc >= 0
        cover += list(range(c0, c0 + nc))
    assert cover == list(range(n))


def test_block_cyclic_owner():
    world = 4
    owners = [zd.owner_of_column(k, world) for k in range(64 * 9)]
    assert owners[0] == 0 and owners[63] == 0 and owners[64] == 1 and owners[64 * 4] == 0 and owners[64 * 7 + 5] == 3
    # every rank owns the same number of blocks (+-1)
    cnt = [sum(1 for b in range(9) if b % world == r) for r in range(world)]
    assert max(cnt) - min(cnt) <= 1


@pytest.mark.parametrize("batch,world", [(1024, 8), (1000, 8), (5, 8), (0, 4), (7, 1), (129, 2)])
def test_batch_shards_partition(batch, world):
    """config 5: contiguous, balanced (+-1) shards that cover the batch exactly"""
    cover, sizes = [], []
    for r in range(world):
        b0, nb = zd.batch_shard(r, world, batch)
        assert 0 <= b0 <= batch and nb >= 0
        cover += list(range(b0, b0 + nb))
        sizes.append(nb)
    assert cover == list(range(batch)) and max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("per,world", [(2048, 8), (4096, 4), (8192, 2), (1024, 8), (1000, 2), (513, 3), (16384, 1), (2050, 2)])
def test_host_pipeline_chunks(per, world, monkeypatch):
    """sub-blocks of the collective host-pointer solve: positive widths that sum to the shard, at most 4, the last one
    (whose download stays exposed) the smallest; small shards are not cut"""
    monkeypatch.delenv("ZQ_DIST_PIPE", raising=False)
    monkeypatch.delenv("ZQ_DIST_CHUNKS", raising=False)
    ch = zd.host_pipeline_chunks(per, world)
    assert 1 <= len(ch) <= 4 and all(c > 0 for c in ch) and sum(ch) == per
    if per < 1024:
        assert ch == [per]
    else:
        assert len(ch) >= 2 and ch[-1] == min(ch) and ch[-1] >= 512
    monkeypatch.setenv("ZQ_DIST_PIPE", "0")
    assert zd.host_pipeline_chunks(per, world) == [per]
    monkeypatch.delenv("ZQ_DIST_PIPE")
    monkeypatch.setenv("ZQ_DIST_CHUNKS", "0,64,32")
    if per > 96:
        assert zd.host_pipeline_chunks(per, world) == [per - 96, 64, 32]


@pytest.mark.parametrize("n,world", [(16384, 8), (16384, 2), (1024, 2), (2050, 2), (4096, 4), (5000, 3), (1030, 8)])
def test_upload_ranges_balanced(n, world):
    """shared upload: the ranges cover [0, n) in order, start on upload-block boundaries, and their lower-triangle
    areas are balanced (within the rounding to 256-column blocks)"""
    b = zd.upload_ranges(n, world)
    assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
    assert all(x % 256 == 0 for x in b[:-1])
    area = [sum(n - c for c in range(b[g], b[g + 1])) for g in range(world)]
    tot = n * (n + 1) // 2
    assert sum(area) == tot
    if n >= 8 * 256 * world:
        assert max(area) <= 1.35 * tot / world
