"""Latency of the drop-in entry at the reference's own sizes: `zquatev_b200(n2, D, ld2, eig)` with HOST pointers in pageable
memory (what a caller of ts::zquatev has), cold (first call of a fresh process: CUDA context, plan allocation) and warm
(same process, later calls; the handle keeps the plan), beside the unmodified reference (oracle/_ref, all host cores) on the
same matrix.  One JSON line per size.  usage: dropin_latency.py [n2 ...]"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, time, json, ctypes, numpy as np
sys.path.insert(0, %r)
from oracle import zquatev_oracle as O
n = int(sys.argv[1])
M = O.gen_testcc(n)[2] if n <= 500 else O.gen_sym(n, 32)
from zquatev_b200 import api
L = api.lib()
ts = []
for it in range(4):
    buf = np.asfortranarray(M).copy(order="F")
    eig = np.zeros(2 * n)
    t0 = time.perf_counter()
    info = L.zquatev_b200(2 * n, buf.ctypes.data, 2 * n, eig.ctypes.data)
    ts.append(time.perf_counter() - t0)
    assert info == 0
w = np.linalg.eigvalsh(M)[0::2] if n <= 1000 else None
dev = float(np.max(np.abs(eig[:n] - w)) / np.abs(w).max()) if w is not None else None
print(json.dumps({"cold_s": ts[0], "warm_s": sorted(ts[1:])[1], "eig_dev_rel_vs_numpy": dev}))
''' % ROOT


def main():
    from oracle import zquatev_oracle as O
    sizes = [int(a) for a in sys.argv[1:]] or [400, 1000, 2000, 4096]
    ref = O.RefLib() if O.RefLib.available() else None
    if ref:
        ref.set_threads(os.cpu_count() or 1)
    for n2 in sizes:
        n = n2 // 2
        out = subprocess.run([sys.executable, "-c", CHILD, str(n)], capture_output=True, text=True, cwd=ROOT)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        rec = json.loads(line[-1]) if line else {"error": out.stderr[-300:]}
        rec["n2"] = n2
        if ref:
            M = O.gen_testcc(n)[2] if n <= 500 else O.gen_sym(n, 32)
            t0 = time.perf_counter()
            ref.zquatev(M)
            rec["reference_s"] = time.perf_counter() - t0
            rec["reference_threads"] = os.cpu_count()
            if "warm_s" in rec:
                rec["warm_speedup"] = rec["reference_s"] / rec["warm_s"]
                rec["cold_speedup"] = rec["reference_s"] / rec["cold_s"]
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
