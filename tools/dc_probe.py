"""D&C probe: solves the tridiagonal produced by the headline workload's statistics (random d, e) at size n
through the K8 test door; used under `ncu --metrics gpu__time_duration.sum -k regex:k_dc` for the per-kernel split."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zquatev_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(0)
# tridiagonal of a reduced random Hermitian matrix: d ~ N(0,1), e_k ~ chi with n-k degrees (Dumitriu-Edelman-like)
d = rng.standard_normal(n)
e = np.sqrt(rng.chisquare(4.0 * np.arange(n - 1, 0, -1)) / 2.0)
dd = torch.from_numpy(d).cuda(); ee = torch.from_numpy(np.concatenate([e, [0.0]])).cuda()
w = torch.zeros(n, dtype=torch.float64, device="cuda"); Z = torch.zeros((n, n), dtype=torch.float64, device="cuda")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    rc = api.lib().zq_test_stedc(n, dd.data_ptr(), ee.data_ptr(), w.data_ptr(), Z.data_ptr())
    torch.cuda.synchronize(); print("stedc n", n, "rc", rc, "wall", time.time() - t0, flush=True)
