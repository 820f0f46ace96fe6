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
    # collective solve with HOST pointers (the reference-shaped entry): shared upload (every rank moves a share of the lower
    # triangles, NVLink carries the rest), back-transformation in column sub-blocks whose exchange + download overlap the next
    # sub-block; host_result 0 (every rank gets all 2n columns) and 1 (rank 0 everything, rank r its own columns); also with
    # the pipeline switched off (gather, then download).  n = 2049: ragged shards (the last rank owns fewer columns).
    import ctypes
    for n, seed in [(700, 6), (2048, 32), (2049, 7)]:
        M = O.gen_sym(n, seed)
        left = torch.from_numpy(np.ascontiguousarray(M[:, :n].T))
        left_dev = left.cuda()
        buf1 = torch.zeros((2 * n, 2 * n), dtype=torch.complex128, device="cuda")
        buf1[:n] = left_dev
        eig1 = torch.zeros(n, dtype=torch.float64, device="cuda")
        i0 = z.zquatev_device(2 * n, buf1.data_ptr(), 2 * n, eig1.data_ptr(), nb=64)          # single-GPU path, this rank
        q1 = G.device_quality(left_dev, buf1, eig1, col_chunk=512)
        nrm = eig1.abs().max().item()
        per = (n + world - 1) // world
        c0 = min(rank * per, n)
        nc = max(0, min(per, n - c0))
        for host_result, pipe in [(1, "1"), (0, "1"), (1, "0")]:
            os.environ["ZQ_DIST_PIPE"] = pipe
            host = torch.full((2 * n, 2 * n), float("nan"), dtype=torch.complex128).pin_memory()
            host[:n] = left
            eig_h = np.zeros(n)
            opt = z.ZqOptions(1, 0, 64, None, 1, 0, 0, 1, host_result)
            info = z.lib().zquatev_b200_ex(2 * n, ctypes.c_void_p(host.data_ptr()), 2 * n, eig_h.ctypes.data, ctypes.byref(opt))
            torch.cuda.synchronize()
            dev_rel = float(np.max(np.abs(eig_h - eig1.cpu().numpy())) / nrm)
            full = host_result == 0 or rank == 0
            ok = info == 0 and i0 == 0 and dev_rel <= 1e-12
            rec = {"transport": "peer", "mode": f"host-pointers host_result={host_result} pipe={pipe}", "n": n, "eig_dev": dev_rel}
            if full:
                hb = host.cuda()
                q = G.device_quality(left_dev, hb, torch.from_numpy(eig_h).cuda(), col_chunk=512)
                ok = ok and q["pairing"] == 0.0 and q["ascending"] and q["residual"] <= max(1.5 * q1["residual"], 0.05) \
                    and q["orthogonality"] <= max(1.5 * q1["orthogonality"], 1.0)
                rec.update({"res": q["residual"], "res_single": q1["residual"], "orth": q["orthogonality"], "orth_single": q1["orthogonality"],
                            "pair": q["pairing"]})
                del hb
            else:
                # own columns and their Kramers partners must be there (finite) and paired exactly
                own = host[c0:c0 + nc]
                part = host[n + c0:n + c0 + nc]
                fin = bool(torch.isfinite(own.view(torch.float64)).all() and torch.isfinite(part.view(torch.float64)).all())
                pair_ok = bool(torch.equal(part[:, :n], -own[:, n:].conj()) and torch.equal(part[:, n:], own[:, :n].conj()))
                ok = ok and fin and pair_ok
                rec.update({"own_cols": [c0, nc], "finite": fin, "paired": pair_ok})
            # the columns this rank holds agree bit for bit with rank 0's copy of them (checksums over torch.distributed)
            cs = torch.tensor([float(host[c].real.sum().item()) + float(host[n + c].imag.sum().item()) for c in (c0, c0 + max(nc - 1, 0))]
                              if nc > 0 else [0.0, 0.0], dtype=torch.float64, device="cuda")
            ref_cs = torch.zeros((world, 2), dtype=torch.float64, device="cuda")
            if rank == 0:
                for r in range(world):
                    r0 = min(r * per, n)
                    rn = max(0, min(per, n - r0))
                    if rn > 0:
                        ref_cs[r] = torch.tensor([float(host[c].real.sum().item()) + float(host[n + c].imag.sum().item()) for c in (r0, r0 + rn - 1)],
                                                 dtype=torch.float64, device="cuda")
            dist.broadcast(ref_cs, 0)
            mine = ref_cs[rank]
            same = bool(torch.equal(mine, cs))
            ok = ok and same
            rec.update({"ok": bool(ok), "cols_match_rank0": same})
            out.append(rec)
            del host
        os.environ.pop("ZQ_DIST_PIPE", None)
    zd.finalize()
    if rank == 0:
        print("DIST_RESULT " + json.dumps(out), flush=True)
    zd.finalize()
    dist.destroy_process_group()
    sys.exit(0 if all(o["ok"] for o in out) else 1)


if __name__ == "__main__":
    main()
