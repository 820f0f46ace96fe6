"""Multi-GPU parity worker (run under torchrun by tests/test_gpu_dist.py / tools): every rank solves
the same matrices with the single-GPU path and with the collective multi-GPU path and compares."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zquatev_b200 as z  # noqa: E402
from oracle import zquatev_oracle as O  # noqa: E402
from zquatev_b200 import dist as zd  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    out = []
    # pass 0: fused peer-memory exchange (CUDA IPC + NVLink stores); pass 1: NCCL per-column collectives
    for transport in ("peer", "nccl"):
      os.environ["ZQ_DIST_NCCL"] = "1" if transport == "nccl" else "0"
      zd.init_from_torch()
      for n, seed in [(3, 1), (64, 2), (65, 3), (130, 4), (333, 5), (700, 6)]:
          M = O.gen_sym(n, seed)
          buf0 = torch.from_numpy(np.asfortranarray(M).T.copy()).cuda()     # column-major memory of M
          res = {}
          for mode in ("single", "dist"):
              buf = buf0.clone()
              eig = torch.zeros(n, dtype=torch.float64, device="cuda")
              info = z.zquatev_device(2 * n, buf.data_ptr(), 2 * n, eig.data_ptr(), nb=64, dist=(mode == "dist"))
              res[mode] = (info, eig.cpu().numpy(), buf.cpu().numpy().T.copy())
          i0, e0, o0 = res["single"]
          i1, e1, o1 = res["dist"]
          nrm = np.abs(e0).max()
          r1, q1, p1 = O.quality(M, o1, e1)
          r0, q0, _ = O.quality(M, o0, e0)
          # every rank must hold the same complete result
          chk = torch.tensor([float(np.abs(o1).sum()), float(e1.sum())], dtype=torch.float64, device="cuda")
          lst = [torch.zeros_like(chk) for _ in range(world)]
          dist.all_gather(lst, chk)
          same = all(torch.equal(lst[0], t) for t in lst)
          ok = (i0 == 0 and i1 == 0 and np.max(np.abs(e0 - e1)) <= 1e-12 * nrm and p1 == 0.0
                and r1 <= max(1.5 * r0, 0.6) and q1 <= max(1.5 * q0, 2.0) and same)
          out.append({"transport": transport, "n": n, "ok": bool(ok), "eig_dev": float(np.max(np.abs(e0 - e1)) / nrm), "res": r1, "res_single": r0,
                      "orth": q1, "orth_single": q0, "pair": p1, "ranks_agree": bool(same)})
    # the paths that only exist at n >= 1024 (3M column-block trailing GEMM) and n >= 2048 (column-split top D&C
    # merge; n % world == 0: in-place all-gather, else broadcast gather), against the golden eigenvalues of the
    # unmodified reference and with the full residual / orthogonality / pairing check on the device
    from tests import gpu_util as G
    gpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sym_large.npz")
    LARGE = dict(np.load(gpath)) if os.path.exists(gpath) else {}
    os.environ["ZQ_DIST_NCCL"] = "0"
    zd.init_from_torch()
    for n, seed in [(1024, 34), (2048, 32), (2050, 33)]:
        if f"eig_{n}_{seed}" not in LARGE:
            continue
        gold, meta = LARGE[f"eig_{n}_{seed}"], LARGE[f"meta_{n}_{seed}"]
        M = O.gen_sym(n, seed)
        left = torch.from_numpy(np.ascontiguousarray(M[:, :n].T)).cuda()
        runs = []
        for rep in range(2):
            buf = torch.full((2 * n, 2 * n), float("nan"), dtype=torch.complex128, device="cuda")
            buf[:n] = left
            eig = torch.zeros(n, dtype=torch.float64, device="cuda")
            info = z.zquatev_device(2 * n, buf.data_ptr(), 2 * n, eig.data_ptr(), nb=64, dist=True)
            runs.append((info, eig, buf))
        info, eig, buf = runs[0]
        repro = bool(torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][2].view(torch.float64), runs[1][2].view(torch.float64)))
        q = G.device_quality(left, buf, eig, col_chunk=512, rank=rank, world=world,
                             reduce=lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM))
        lst = [torch.zeros_like(eig) for _ in range(world)]
        dist.all_gather(lst, eig)
        same = all(torch.equal(lst[0], t) for t in lst)
        dev_gold = float(np.max(np.abs(eig.cpu().numpy() - gold)) / meta[5])
        ok = (info == 0 and same and repro and dev_gold <= 1e-12 and q["pairing"] == 0.0 and q["ascending"]
              and q["residual"] <= meta[1] and q["orthogonality"] <= meta[2])
        out.append({"transport": "peer", "n": n, "ok": bool(ok), "eig_dev_vs_reference": dev_gold, "res": q["residual"], "res_reference": float(meta[1]),
                    "orth": q["orthogonality"], "orth_reference": float(meta[2]), "pair": q["pairing"], "ranks_agree": bool(same),
                    "bitwise_reproducible": repro})
    # values-only collective solve (BASELINE config 4 shape): 1-D block-cyclic reduction + bisection sharded by
    # eigenvalue index ranges with one all-gather of n doubles
    for n, seed in [(65, 3), (333, 5), (1030, 8)]:
        M = O.gen_sym(n, seed)
        buf0 = torch.from_numpy(np.asfortranarray(M).T.copy()).cuda()
        e_full = torch.zeros(n, dtype=torch.float64, device="cuda")
        b1 = buf0.clone()
        i0 = z.zquatev_device(2 * n, b1.data_ptr(), 2 * n, e_full.data_ptr(), nb=64)
        e_val = torch.zeros(n, dtype=torch.float64, device="cuda")
        b2 = buf0.clone()
        i1 = z.zquatev_device(2 * n, b2.data_ptr(), 2 * n, e_val.data_ptr(), jobz=0, nb=64, dist=True)
        nrm = e_full.abs().max().item()
        dev_rel = (e_val - e_full).abs().max().item() / nrm
        lst = [torch.zeros_like(e_val) for _ in range(world)]
        dist.all_gather(lst, e_val)
        same = all(torch.equal(lst[0], t) for t in lst)
        ok = i0 == 0 and i1 == 0 and dev_rel <= 1e-12 and same and bool(torch.all(e_val[1:] >= e_val[:-1]))
        out.append({"transport": "peer", "mode": "values-only", "n": n, "ok": bool(ok), "eig_dev": dev_rel, "ranks_agree": bool(same)})
    zd.finalize()
    if rank == 0:
        print("DIST_RESULT " + json.dumps(out), flush=True)
    zd.finalize()
    dist.destroy_process_group()
    sys.exit(0 if all(o["ok"] for o in out) else 1)


if __name__ == "__main__":
    main()
