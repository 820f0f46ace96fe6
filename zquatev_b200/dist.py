"""Multi-GPU plumbing: one process per GPU, torch.distributed carries the 128-byte NCCL unique id,
the C library owns the communicator used inside the solver."""
from __future__ import annotations

import ctypes

from . import api


def column_block(rank: int, world: int, n: int):
    """Eigenvector column block [col0, col0 + ncols) that `rank` back-transforms (mirrors
    solve_device in csrc/solver.cu): ceil(n / world) columns per rank, the last ranks may get fewer."""
    per = (n + world - 1) // world
    col0 = min(rank * per, n)
    return col0, max(0, min(per, n - col0))


def batch_shard(rank: int, world: int, batch: int):
    """Problems [b0, b0 + nb) of a batch of independent small matrices that `rank` solves (BASELINE
    config 5: one batch shard per GPU, no collective on the data path): contiguous shards whose sizes
    differ by at most one, so the caller's arrays are sliced without copies."""
    base, extra = divmod(batch, world)
    b0 = rank * base + min(rank, extra)
    return b0, base + (1 if rank < extra else 0)


def host_pipeline_chunks(per: int, world: int):
    """Widths of the sub-blocks in which every rank back-transforms its `per` eigenvector columns in the collective
    host-pointer solve (finished sub-blocks are exchanged and downloaded while the next is computed); the planner lives
    in csrc/solver.cu (dist_sink_chunks), this is its door."""
    out = (ctypes.c_int * 4)()
    cnt = api.lib().zq_test_dist_chunks(per, world, out)
    return [out[i] for i in range(cnt)]


def upload_ranges(n: int, world: int):
    """Column ranges [b[g], b[g+1]) whose lower triangles rank g uploads in the collective host-pointer solve (equal
    triangle areas; csrc/solver.cu upload_bounds)."""
    b = (ctypes.c_int * (world + 1))()
    rc = api.lib().zq_test_upload_bounds(n, world, b)
    if rc != 0:
        raise ValueError("world out of range")
    return list(b)


def owner_of_column(k: int, world: int, nb: int = 64) -> int:
    """Rank that owns column k of (D; E) in the 1-D block-cyclic layout of the reduction."""
    return (k // nb) % world


def exchange_unique_id(make_id, group=None):
    """rank 0 calls `make_id()` -> 128 bytes; everyone receives them through torch.distributed
    (works with the nccl and the gloo backend)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        buf = torch.tensor(list(raw), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        dev = buf.cuda()
        dist.broadcast(dev, 0, group=group)
        buf = dev.cpu()
    else:
        dist.broadcast(buf, 0, group=group)
    return bytes(buf.tolist())


def _make_nccl_id() -> bytes:
    raw = (ctypes.c_ubyte * 128)()
    rc = api.lib().zquatev_b200_dist_unique_id(raw)
    if rc != 0:
        raise RuntimeError(f"zquatev_b200_dist_unique_id failed: {rc}")
    return bytes(raw)


def init_from_torch(group=None):
    """Creates the solver's NCCL communicator on the current CUDA device for all ranks of the
    (already initialised) torch.distributed process group.  Returns (rank, world)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    raw = exchange_unique_id(_make_nccl_id, group)
    buf = (ctypes.c_ubyte * 128).from_buffer_copy(raw)
    rc = api.lib().zquatev_b200_dist_init(rank, world, buf)
    if rc != 0:
        raise RuntimeError(f"zquatev_b200_dist_init failed: {rc}")
    return rank, world


def finalize():
    api.lib().zquatev_b200_dist_finalize()


def zquatev_batched_sharded(D, eig, group=None):
    """Config 5 across the ranks of a torch.distributed group: every rank holds the same (batch, n2, n2) / (batch, n)
    host arrays and solves only its own shard in place through `zquatev_batched` (independent problems: no data-path
    collective).  Returns (b0, nb, info) of the local shard; gathering the shards is the caller's choice."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    b0, nb = batch_shard(rank, world, D.shape[0])
    info = api.zquatev_batched(D[b0:b0 + nb], eig[b0:b0 + nb]) if nb > 0 else None
    return b0, nb, info
