"""BASELINE config 4: 2n = 65536 eigenvalues only (the 34 GB left half resident in HBM) on 1 GPU (`python
tools/config4_dist.py`) or as one collective solve over the GPUs of a box (`torchrun --nproc-per-node G ...`): D and E 1-D
block-cyclic, bisection sharded by eigenvalue index ranges.  Checks: trace, sum of squares, ascending order, identical
eigenvalues on every rank, and agreement with the single-GPU eigenvalues when profiles/r02_config4_eig_1gpu.npy exists
(written by the single-GPU run with --save)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zquatev_b200 as z  # noqa: E402
from bench import make_input  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 32768
save = "--save" in sys.argv
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=600))
    from zquatev_b200 import dist as zd
    zd.init_from_torch()
left0 = make_input(n, dev)                     # identical on every rank (seeded)
idx = torch.arange(n, device=dev)
tr = left0[idx, idx].real.sum().item()
ss = 0.0                                       # ||M||_F^2 / 2 = ||D||_F^2 + ||E||_F^2 from the lower triangles
chunk = max(1, (1 << 26) // (2 * n))
for c0 in range(0, n, chunk):
    c1 = min(n, c0 + chunk)
    blk = left0[c0:c1]
    low = torch.arange(n, device=dev)[None, :] > torch.arange(c0, c1, device=dev)[:, None]
    ss += 2.0 * ((blk[:, :n].abs() ** 2) * low).sum().item() + 2.0 * ((blk[:, n:].abs() ** 2) * low).sum().item()
ss += (left0[idx, idx].real ** 2).sum().item()
work = torch.empty_like(left0)
eig = torch.zeros(n, dtype=torch.float64, device=dev)
times = []
for it in range(2):                            # first = warm-up (plan, peer buffers)
    work.copy_(left0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = z.zquatev_device(2 * n, work.data_ptr(), 2 * n, eig.data_ptr(), jobz=0, nb=64, dist=world > 1)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
ph = z.last_phases()
tsec = torch.tensor([times[-1]], dtype=torch.float64, device=dev)
same = True
if world > 1:
    dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
    lst = [torch.zeros_like(eig) for _ in range(world)]
    dist.all_gather(lst, eig)
    same = all(torch.equal(lst[0], t) for t in lst)
eh = eig.cpu().numpy()
ref_path = os.path.join(ROOT, "profiles", "r02_config4_eig_1gpu.npy")
vs1 = None
if world > 1 and os.path.exists(ref_path) and n == 32768:
    e1 = np.load(ref_path)
    vs1 = float(np.max(np.abs(e1 - eh)) / np.abs(e1).max())
if rank == 0:
    if save and world == 1:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", "r02_config4_eig_1gpu.npy"), eh)
    dt = tsec.item()
    print(json.dumps({"config": 4, "gpus": world, "n2": 2 * n, "info": info, "seconds": dt, "seconds_first_call": times[0], "phases_ms": ph,
                      "trace_err": abs(float(eh.sum()) - tr), "sumsq_rel_err": abs(float((eh ** 2).sum()) - ss) / ss,
                      "ascending": bool(np.all(eh[1:] >= eh[:-1])), "ranks_identical": bool(same), "eig_vs_1gpu_rel": vs1,
                      "tflops_values_only_canonical": 64.0 / 3.0 * n ** 3 / dt * 1e-12}), flush=True)
if world > 1:
    zd.finalize()
    dist.destroy_process_group()
