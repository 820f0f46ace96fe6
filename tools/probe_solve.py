"""Development probe: one device-resident solve at half-dimension n (optionally only an eigenvector column block,
as a rank of the multi-GPU solve would do) and the phase split.  usage: probe_solve.py n [ncols] [reps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zquatev_b200 as z  # noqa: E402
from bench import make_input  # noqa: E402

n = int(sys.argv[1])
ncols = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
left0 = make_input(n, dev)
work = torch.empty((2 * n, 2 * n), dtype=torch.complex128, device=dev)
eig = torch.zeros(n, dtype=torch.float64, device=dev)
# warm the kernels / attributes on a small problem of the same code path (n >= 1024 -> 3M GEMM)
ws = make_input(1024, dev)
ww = torch.empty((2048, 2048), dtype=torch.complex128, device=dev)
ww[:1024].copy_(ws)
z.zquatev_device(2048, ww.data_ptr(), 2048, eig.data_ptr())
for _ in range(reps):
    work[:n].copy_(left0)
    info = z.zquatev_device(2 * n, work.data_ptr(), 2 * n, eig.data_ptr(), col0=0, ncols=ncols)
    ph = z.last_phases()
    tr = torch.diagonal(left0[:, :n]).real.sum().item()
    print(json.dumps({"n": n, "ncols": ncols or n, "info": info, "env": {k: v for k, v in os.environ.items() if k.startswith("ZQ_")},
                      "phases_ms": {k: round(v, 2) for k, v in ph.items()}, "trace_err": abs(eig.sum().item() - tr)}), flush=True)
