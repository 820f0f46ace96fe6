"""Development probe for the collective (multi-GPU) solve, run under torchrun: A/B of the knobs that only exist at N > 1.
  device-resident : ZQ_DIST_EARLY_PUSH (look-ahead of the panel exchange), ZQ_DC_SPLIT_LEVELS (lower D&C levels split)
  host pointers   : ZQ_DIST_PIPE / ZQ_DIST_UPLOAD / ZQ_DIST_CHUNKS (pipelined delivery, shared upload)
usage: torchrun ... tools/dist_probe.py [n]   -> JSON lines on rank 0"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zquatev_b200 as z  # noqa: E402
from bench import make_input  # noqa: E402
from zquatev_b200 import dist as zd  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    short = len(sys.argv) > 2 and sys.argv[2] in ("short", "hostonly")   # fewer A/B legs (large rank counts are charged per GPU)
    hostonly = len(sys.argv) > 2 and sys.argv[2] == "hostonly"           # one device-resident solve (warm-up + reference eigenvalues)
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    zd.init_from_torch()
    n2 = 2 * n
    left0 = make_input(n, dev)
    work = torch.empty((n2, n2), dtype=torch.complex128, device=dev)
    eig = torch.zeros(n, dtype=torch.float64, device=dev)

    def setenv(env):
        for k in ("ZQ_DIST_EARLY_PUSH", "ZQ_DC_SPLIT_LEVELS", "ZQ_DIST_PIPE", "ZQ_DIST_UPLOAD", "ZQ_DIST_CHUNKS"):
            os.environ.pop(k, None)
        os.environ.update(env)

    def sync():
        dist.barrier()
        torch.cuda.synchronize()

    ref_eig = None
    for env in ([{}] if hostonly else [{}, {}, {"ZQ_DIST_EARLY_PUSH": "0", "ZQ_DC_SPLIT_LEVELS": "0"}] if short else
                [{}, {}, {"ZQ_DIST_EARLY_PUSH": "0"}, {"ZQ_DC_SPLIT_LEVELS": "0"}]):
        setenv(env)
        work[:n].copy_(left0)
        sync()
        t0 = time.perf_counter()
        info = z.zquatev_device(n2, work.data_ptr(), n2, eig.data_ptr(), dist=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if ref_eig is None:
            ref_eig = eig.clone()
        same = bool(torch.equal(ref_eig, eig))
        if rank == 0:
            print(json.dumps({"mode": "device", "n": n, "world": world, "env": env, "info": info, "wall_s_max": t.item(),
                              "eig_identical_to_first": same, "phases_ms": {k: round(v, 2) for k, v in z.last_phases().items()},
                              "gather_ms": z.last_gather_ms()}), flush=True)
    del work
    torch.cuda.empty_cache()
    host = torch.empty((n2, n2), dtype=torch.complex128, pin_memory=True)
    host0 = left0.cpu()
    eig_h = np.zeros(n2)
    for env in ([{}, {}, {"ZQ_DIST_PIPE": "0", "ZQ_DIST_UPLOAD": "0"}] if short else
                [{}, {}, {"ZQ_DIST_PIPE": "0", "ZQ_DIST_UPLOAD": "0"}, {"ZQ_DIST_PIPE": "0"}, {"ZQ_DIST_UPLOAD": "0"}]):
        setenv(env)
        host[:n].copy_(host0)
        sync()
        t0 = time.perf_counter()
        opt = z.ZqOptions(1, 0, 0, None, 1, 0, 0, 1, 1)
        info = z.lib().zquatev_b200_ex(n2, ctypes.c_void_p(host.data_ptr()), n2, eig_h.ctypes.data, ctypes.byref(opt))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dmax = float(np.max(np.abs(eig_h[:n] - ref_eig.cpu().numpy())))
        if rank == 0:
            print(json.dumps({"mode": "host", "n": n, "world": world, "env": env, "info": info, "e2e_s_max": t.item(),
                              "eig_max_abs_diff_vs_device": dmax, "chunks": zd.host_pipeline_chunks((n + world - 1) // world, world),
                              "phases_ms": {k: round(v, 2) for k, v in z.last_phases().items()}}), flush=True)
    setenv({})
    zd.finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
