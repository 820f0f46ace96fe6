"""Prints the north_star quality numbers of the CUDA path next to the reference's (golden)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import zquatev_oracle as O
from tests import gpu_util as G
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
T = json.load(open(os.path.join(GOLD, "testcc_eigs.json")))
S = json.load(open(os.path.join(GOLD, "sym_eigs.json")))
rows = []
for n in [1, 2, 3, 21, 22, 23, 64, 200, 500]:
    _, _, C = O.gen_testcc(n)
    eig, out, info = G.solve_host(C)
    g = T[str(n)]
    res, orth, pair = O.quality(C, out, eig[:n])
    rows.append({"case": f"testcc n={n}", "info": info, "eig_dev_over_norm": float(np.max(np.abs(eig[:n] - np.array(g["eig"]))) / g["two_norm"]),
                 "res": res, "res_ref": g["residual"], "orth": orth, "orth_ref": g["orthogonality"], "pair": pair})
for key in S:
    n, seed = (int(x) for x in key.split("_"))
    M = O.gen_sym(n, seed)
    eig, out, info = G.solve_host(M)
    g = S[key]
    res, orth, pair = O.quality(M, out, eig[:n])
    rows.append({"case": f"sym {key}", "info": info, "eig_dev_over_norm": float(np.max(np.abs(eig[:n] - np.array(g["eig"]))) / g["two_norm"]),
                 "res": res, "res_ref": g["residual"], "orth": orth, "orth_ref": g["orthogonality"], "pair": pair})
if O.RefLib.available():
    ref = O.RefLib()
    for n in [1000]:
        M = O.gen_sym(n, 32)
        er, outr, _ = ref.zquatev(M)
        eig, out, info = G.solve_host(M)
        rr, orr, _ = O.quality(M, outr, er)
        res, orth, pair = O.quality(M, out, eig[:n])
        rows.append({"case": f"sym {n}_32 vs live reference", "info": info, "eig_dev_over_norm": float(np.max(np.abs(eig[:n] - er)) / np.abs(er).max()),
                     "res": res, "res_ref": rr, "orth": orth, "orth_ref": orr, "pair": pair})
for r in rows:
    print(json.dumps(r))
