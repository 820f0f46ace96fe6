"""Golden eigenvalues of the UNMODIFIED reference (oracle/_ref) for the LARGE G_sym cases, stored compactly as
tests/golden/sym_large.npz (run in the build container: `make -C oracle && python tests/golden/make_golden_large.py`).

  G_sym(4096, 32)  -- BASELINE.json configs[1] (2n = 8192), the single-GPU parity case
  G_sym(2048, 32), G_sym(2050, 33) -- the multi-GPU parity cases (n divisible / not divisible by the world size:
                      column-split top D&C merge with the in-place all-gather vs the broadcast gather)
  G_sym(1024, 34)  -- first size of the 3M trailing-update path

Per case: all n eigenvalues returned by ts::zquatev (zquatev.cc:42-100), its info, and the reference's own
north_star quality numbers (residual, orthogonality, pairing; test.cc:104-112 restated in oracle.quality) so the GPU
tests can assert "at or below the reference's".  zheev (test.cc:84-95) is skipped at these sizes (2n = 8192 takes
~40 min on 8 cores); the reference's own eigenvalues are the pin.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import zquatev_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(1024, 34), (2048, 32), (2050, 33), (4096, 32)]


def main():
    ref = O.RefLib()
    ref.set_threads(os.cpu_count() or 1)
    out = {}
    path = os.path.join(HERE, "sym_large.npz")
    for n, seed in CASES:
        M = O.gen_sym(n, seed)
        t0 = time.perf_counter()
        eig, vec, info = ref.zquatev(M)
        dt = time.perf_counter() - t0
        res, orth, pair = O.quality(M, vec, eig)
        out[f"eig_{n}_{seed}"] = eig
        out[f"meta_{n}_{seed}"] = np.array([info, res, orth, pair, np.linalg.norm(M), np.abs(eig).max(), dt])
        print(f"sym {n} {seed}: info={info} res={res:.4f} orth={orth:.4f} pair={pair} eig0={eig[0]:.12f} ref_seconds={dt:.1f}", flush=True)
        np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
