"""BASELINE configs 4 and 5 (parity/property checks + timings, not bench lines):
 config 4: 2n=65536 eigenvalues only (34 GB left half resident in HBM, K1-K4 + K9 bisection), checked by trace
           and sum of squares;  config 5: batch of 2n=512 problems through zquatev_b200_batched, checked against
           numpy on a sample and timed."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zquatev_b200 as z
from oracle import zquatev_oracle as O

which = sys.argv[1] if len(sys.argv) > 1 else "5"
if which == "4":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(32)
    left = torch.empty((n, 2 * n), dtype=torch.complex128, device=dev)     # column-major (A;B), 16*2n*n bytes
    chunk = max(1, (1 << 26) // (2 * n))
    tr = 0.0; ss = 0.0
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        X = torch.rand((c1 - c0, 2 * n, 2), dtype=torch.float64, device=dev, generator=g) - 0.5
        left[c0:c1] = torch.view_as_complex(X)
    idx = torch.arange(n, device=dev)
    left[idx, idx] = left[idx, idx].real.to(torch.complex128)
    tr = left[idx, idx].real.sum().item()
    # ||M||_F^2 = 2 (||D||^2 + ||E||^2) from the lower triangles the solver reads
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        blk = left[c0:c1]
        cols = torch.arange(c0, c1, device=dev)[:, None]
        rows = torch.arange(n, device=dev)[None, :]
        low = rows > cols
        d2 = (blk[:, :n].abs() ** 2)
        e2 = (blk[:, n:].abs() ** 2)
        ss += 2.0 * (d2 * low).sum().item() + 2.0 * (e2 * low).sum().item()
    ss += (left[idx, idx].real ** 2).sum().item()
    fro2 = 2.0 * ss
    eig = torch.zeros(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize(); t0 = time.time()
    info = z.zquatev_device(2 * n, left.data_ptr(), 2 * n, eig.data_ptr(), jobz=0)
    torch.cuda.synchronize(); dt = time.time() - t0
    ph = z.last_phases()
    res = {"config": 4, "n2": 2 * n, "info": info, "seconds": dt, "phases_ms": ph, "trace_err": abs(eig.sum().item() - tr),
           "sumsq_rel_err": abs((eig ** 2).sum().item() - 0.5 * fro2) / (0.5 * fro2), "ascending": bool(torch.all(eig[1:] >= eig[:-1])),
           "tflops_values_only": 64.0 / 3.0 * n ** 3 / dt * 1e-12}
    print(json.dumps(res), flush=True)
else:
    n = 256
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    rng = np.random.default_rng(5)
    Ms = [O.gen_sym(n, 1000 + b) for b in range(min(batch, 4))]
    D = np.empty((batch, 2 * n, 2 * n), dtype=np.complex128)
    for b in range(batch):
        D[b] = np.asfortranarray(Ms[b % len(Ms)]).T
    eig = np.zeros((batch, n))
    D2 = D.copy()
    t0 = time.time(); info = z.zquatev_batched(D2, eig); dt0 = time.time() - t0      # includes plan creation
    D2 = D.copy()
    t0 = time.time(); info = z.zquatev_batched(D2, eig); dt = time.time() - t0
    ok = bool(np.all(info == 0))
    worst = 0.0; wq = (0.0, 0.0, 0.0)
    for b in [0, 1, 2, 3, batch - 1]:
        M = Ms[b % len(Ms)]
        wr = np.linalg.eigvalsh(M)[0::2]
        worst = max(worst, float(np.max(np.abs(eig[b] - wr)) / np.abs(wr).max()))
        q = O.quality(M, D2[b].T, eig[b])
        wq = tuple(max(a, c) for a, c in zip(wq, q))
    # one-at-a-time loop through the single-problem entry for comparison
    D3 = D.copy(); e1 = np.zeros(2 * n)
    t0 = time.time()
    for b in range(min(batch, 32)):
        z.zquatev(2 * n, D3[b], 2 * n, e1)
    dloop = (time.time() - t0) / min(batch, 32)
    print(json.dumps({"config": 5, "batch": batch, "n2": 2 * n, "all_info_zero": ok, "seconds_first_call": dt0, "seconds": dt,
                      "ms_per_matrix": dt / batch * 1e3, "ms_per_matrix_single_entry_loop": dloop * 1e3, "max_eig_dev_rel": worst,
                      "max_res_orth_pair": wq, "graph_launches_eager_solves": z.batched_stats(), "gflops_canonical": 164.0 / 3.0 * n ** 3 * batch / dt * 1e-9}), flush=True)
