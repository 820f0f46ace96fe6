"""Generates tests/golden/*.json from the UNMODIFIED reference built in oracle/_ref
(run in the build container where /root/reference exists: `make -C oracle && python tests/golden/make_golden.py`).

testcc_eigs.json : for the test.cc:58-78 matrix G_testcc(n): all n eigenvalues returned by the reference
                   ts::zquatev, the reference's own quality numbers (north_star metrics) and the two numbers
                   test.cc prints (test.cc:104-112).
sym_eigs.json    : same for G_sym(n, seed) (oracle.zquatev_oracle.gen_sym).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import zquatev_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def record(M, ref):
    eig, out, info = ref.zquatev(M)
    ez, _, _ = ref.zheev(M)
    res, orth, pair = O.quality(M, out, eig)
    err, maxdev = O.testcc_checks(M, out, eig, ez)
    return {"info": info, "eig": [float(x) for x in eig], "residual": res, "orthogonality": orth, "pairing": pair,
            "testcc_error": err, "testcc_maxdev": maxdev, "fro_norm": float(np.linalg.norm(M)),
            "two_norm": float(np.max(np.abs(ez)))}


def main():
    ref = O.RefLib()
    out = {}
    for n in [1, 2, 3, 21, 22, 23, 64, 200, 500]:
        _, _, C = O.gen_testcc(n)
        out[str(n)] = record(C, ref)
        print("testcc", n, out[str(n)]["eig"][0], out[str(n)]["residual"], out[str(n)]["orthogonality"])
    json.dump(out, open(os.path.join(HERE, "testcc_eigs.json"), "w"))
    out = {}
    for n, seed in [(5, 32), (33, 33), (100, 34), (257, 32), (256, 1000)]:   # (256, 1000): first problem of the config-5 batch
        M = O.gen_sym(n, seed)
        out[f"{n}_{seed}"] = record(M, ref)
        print("sym", n, seed, out[f"{n}_{seed}"]["eig"][0], out[f"{n}_{seed}"]["residual"])
    json.dump(out, open(os.path.join(HERE, "sym_eigs.json"), "w"))


if __name__ == "__main__":
    main()
