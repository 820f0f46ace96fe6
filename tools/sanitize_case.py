"""One small solve per code path for `compute-sanitizer` (memcheck / racecheck / initcheck are 10-100x slower than a plain run):
K5 one-CTA reduction (n = 64), the column chain with K1 runs and the 4-product GEMMs (n = 130, ZQ_SMALL_N=0), values-only
bisection, and -- with ZQ_QGEMM_MIN lowered by the caller -- nothing else: the n >= 1024 kernels are exercised by the tests."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import zquatev_oracle as O  # noqa: E402
from tests import gpu_util as G  # noqa: E402

for n, small, jobz in [(64, "256", 1), (130, "0", 1), (70, "0", 0)]:
    os.environ["ZQ_SMALL_N"] = small
    M = O.gen_sym(n, 3)
    eig, out, info = G.solve_host(M, jobz=jobz)
    w = np.linalg.eigvalsh(M)[0::2]
    print(n, small, jobz, info, float(np.max(np.abs(eig[:n] - w))))
