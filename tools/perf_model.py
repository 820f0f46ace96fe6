"""Phase-time model of the one-stage solver (documentation aid, not a measurement): per-kernel constants taken from
the round-1 measurements (profiles/r01_summary.md) are summed over the n-1 columns / n/nb panels of a 2n x 2n
problem on G GPUs.  usage: python tools/perf_model.py [n] -- prints predicted vs measured where a measurement exists."""
import json
import sys

NB = 64
HBM = 5.75e12          # B/s K1 sustains on large launches (ncu: 5.69-5.87 TB/s)
K1_T0 = 14e-6          # fixed ramp/tail per K1 launch (fit to the 4.51-4.59 s K1 total of the 1-GPU step)
CHAIN_C0 = 16e-6       # col_update + reflector + reduce_correct at m -> 0 with programmatic dependent launch
CHAIN_C1 = 2.8e-9      # ... + seconds per row (62 us at m = 15 600 in the ncu window)
def px_extra(G):       # multi-GPU: extra launches + two NVLink exchanges per column; ~2 us at G = 2 (measured, session 3),
    return 0.0 if G == 1 else 2e-6 + 5e-6 * (G - 2)   # ~40 us at G = 8 before programmatic launch (session 2): linear guess between
K4_RATE = 37.1e12      # canonical flop/s of the trailing update inside a step
K6_RATE = 39.9e12      # canonical flop/s of the back-transformation inside a step
DC_1 = 0.273           # s at n = 16384, replicated; the top-level merge GEMM (~40 %) is column-split over the ranks
GATHER_BW = 300e9      # B/s per rank assumed for the NCCL all-gather of the eigenvector shards


def model(n, G):
    k1 = chain = k4 = 0.0
    for k in range(n - 1):
        m = n - 1 - k
        k1 += K1_T0 + 16.0 * m * m / G / HBM
        chain += CHAIN_C0 + CHAIN_C1 * m + px_extra(G)
    for j0 in range(0, n - 1, NB):
        kb = min(NB, n - 1 - j0)
        m = n - (j0 + kb)
        k4 += 32.0 * m * m * kb / G / K4_RATE
    dc = DC_1 * (n / 16384.0) ** 3 * (0.6 + 0.4 / G)
    bt = 32.0 * n ** 3 / G / K6_RATE * (1.0 if G == 1 else 1.0 + 0.02 * G)     # narrower column shards run a little slower
    gather = 0.0 if G == 1 else (G - 1) / G * 2 * 16.0 * n * (2 * n) / GATHER_BW
    tri = k1 + chain + k4
    return {"k1": k1, "chain": chain, "k4": k4, "tridiag": tri, "dc": dc, "backtransform": bt + gather, "gather": gather,
            "total": tri + dc + bt + gather}


MEASURED = {(16384, 1): {"tridiag": 6.322, "dc": 0.273, "backtransform": 3.524, "total": 10.118, "src": "r01_bench_1gpu.json"},
            (16384, 2): {"tridiag": 3.564, "dc": 0.214, "backtransform": 1.790, "total": 5.572, "src": "r01_bench_2gpu_pdl.jsonl"},
            (16384, 8): {"tridiag": 2.183, "dc": 0.267, "backtransform": 1.149, "total": 3.603, "src": "r01_bench_8gpu.json (session 2 code)"},
            (4096, 1): {"tridiag": 0.188, "dc": 0.013, "backtransform": 0.0605, "total": 0.262, "src": "r01_bench_2n8192.json"}}

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    for G in (1, 2, 4, 8):
        r = {k: round(v, 3) for k, v in model(n, G).items()}
        print(json.dumps({"n": n, "gpus": G, "model_s": r, "measured_s": MEASURED.get((n, G))}), flush=True)
