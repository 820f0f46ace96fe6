"""BASELINE config 5 across the GPUs of one box: `torchrun --nproc-per-node G tools/config5_dist.py [batch]` -- every
rank builds the same synthetic batch of 2n = 512 problems, solves its contiguous shard through the host-pointer batched
entry (no data-path collective) and rank 0 prints the whole-job throughput (max over ranks of the shard time)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zquatev_b200 as z  # noqa: E402
from oracle import zquatev_oracle as O  # noqa: E402  (checker only)
from zquatev_b200 import dist as zd  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = 256
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Ms = [O.gen_sym(n, 1000 + b) for b in range(4)]
D = np.empty((batch, 2 * n, 2 * n), dtype=np.complex128)
for b in range(batch):
    D[b] = np.asfortranarray(Ms[b % 4]).T
eig = np.zeros((batch, n))
zd.zquatev_batched_sharded(D.copy(), eig.copy())          # warm-up: lanes, graphs
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
b0, nb, info = zd.zquatev_batched_sharded(D, eig)
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
ok = bool(info is None or np.all(info == 0))
wr = np.linalg.eigvalsh(Ms[b0 % 4])[0::2] if nb else None
dev = float(np.max(np.abs(eig[b0] - wr)) / np.abs(wr).max()) if nb else 0.0
if int(os.environ.get("RANK", "0")) == 0:
    print(json.dumps({"config": 5, "gpus": world, "batch": batch, "n2": 2 * n, "seconds": dt.item(), "matrices_per_s": batch / dt.item(),
                      "shard_rank0": [b0, nb], "all_info_zero": ok, "eig_dev_rel_first": dev}), flush=True)
if world > 1:
    dist.destroy_process_group()
