#!/usr/bin/env python
"""bench.py -- headline benchmark of the zquatev hot path (contract: README of the build task).

metric  : time-to-solution (s) of the full eigendecomposition (eigenvalues + symmetry-adapted
          eigenvectors) of a 2n x 2n quaternionic Hermitian matrix, BASELINE.json's metric
          ("2n=32768 eigvals+vecs time-to-solution"); lower is better.  The canonical FP64 rate
          F_total / time (F_total = 164/3 n^3) is reported next to it.
step    : one complete solve of one synthetic matrix (G_sym: uniform [-1/2,1/2) entries, seed 32).
value   : device-resident (input already in HBM, result left in HBM), max over ranks.
e2e     : the same solve through the reference-facing C ABI `zquatev_b200_ex` with HOST (pinned)
          buffers: H2D of the left half and D2H of all 2n columns inside the timed region.
N > 1   : strong scaling (the problem is fixed).  One collective solve: D and E distributed 1-D
          block-cyclic by 64-column blocks; per panel the owner pushes its 64 columns to every rank and per
          column the partial mat-vecs are exchanged, both as peer-memory stores fused into the panel kernels
          (CUDA IPC over NVLink; NCCL collectives as fallback); trailing update on owned blocks only;
          the top three levels of the tridiagonal D&C and the back-transformation are split by eigenvector
          columns; result gathered on every rank (device-resident) or, with host pointers, exchanged and
          downloaded sub-block by sub-block while the next sub-block is back-transformed.
quality : after the timed loop (outside it) the result of the LAST timed solve is checked on the device at the
          full size, on every N: residual ||MV-VL||_F/(N||M||_F eps), orthogonality ||V^H V-I||_F/(N eps), exact
          quaternion pairing, trace, sum of squares, ascending order (the checks of test.cc:104-112); for N > 1
          also that all ranks hold bit-identical eigenvalues and that they agree with a single-GPU solve of the
          same matrix within 1e-12 ||A||.  The run FAILS (assert) if res >= 0.5, orth >= 1.5 or pairing != 0.
--impl reference : the UNMODIFIED reference ts::zquatev built in oracle/_ref on the host cores.  Each step times a
          bounded sample (2n = 2048); once per run ts::zquatev AND LAPACK zheev on the 2n matrix (test.cc:84-102)
          are timed at n = 500, 1024, 2048 (zheev: 500, 1024), power laws a*n^p are fitted and the value is
          EXTRAPOLATED to the workload with the fitted exponent (a real 2n=32768 CPU run takes
          hours).  The same three-point measurement is the `cpu_baseline` of the GPU arm at N = 1.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "2n=%d eigvals+vecs time-to-solution"
# executed / canonical GEMM flops of the n >= 1024 kernels: 0.5 for the quaternion 8-product GEMM (8 real products where the
# canonical count -- 4 complex products of 4 real products -- has 16), 0.75 for the stacked complex 3M form, 1 for 4 products
GEMM_EXEC = (0.5 if os.environ.get("ZQ_QGEMM", "1") != "0" else (0.75 if os.environ.get("ZQ_GEMM_3M", "1") != "0" else 1.0))
GEMM_NAME = ("k_qgemm8 (quaternion GEMM, 8 real DMMA products per quaternion product)" if GEMM_EXEC == 0.5 else
             "k_zgemm_3m (stacked complex GEMM, 3 real products per complex product)" if GEMM_EXEC == 0.75 else "k_zgemm_mma (4 products)")
REF_SAMPLE_N = 1024          # 2n = 2048 reference run per step (~5-10 s on 16 cores)
REF_FIT_ZQ = (500, 1024, 2048)   # sizes at which the reference is timed for the power-law fit
REF_FIT_ZHEEV = (500, 1024)      # zheev(2n) at 2n = 4096 alone is ~5 min on 8 cores: left out of the bounded sample
WORKLOAD = ("2n=%d quaternionic Hermitian eigendecomposition (values+vectors), G_sym seed 32 "
            "[GPU arm: measured at this size; CPU reference arm: EXTRAPOLATED to this size from measured "
            "2n<=4096 samples with a fitted power law a*n^p]")


def bench_config(n2, nb):
    """identical in both arms (the driver compares them)"""
    return {"workload": WORKLOAD % n2, "n2": n2, "nb": nb or 64,
            "l2": "inputs (16*n2*n B) larger than L2; fresh copy of the input every step"}


def fit_power(ns, ts):
    """least-squares a*n^p on log-log; returns (a, p)"""
    import math
    xs, ys = [math.log(x) for x in ns], [math.log(y) for y in ts]
    k = len(xs)
    if k < 2:
        return (ts[0] / ns[0] ** 3, 3.0)
    mx, my = sum(xs) / k, sum(ys) / k
    p = sum((x - mx) * (y - my) for x, y in zip(xs, ys)) / sum((x - mx) ** 2 for x in xs)
    return (math.exp(my - p * mx), p)


def cpu_reference_points(ref, O, n_target, known=None):
    """times ts::zquatev (reference, oracle/_ref) and LAPACK zheev(2n) (as test.cc:84-102 does) on the host cores at
    the BASELINE.md 3 sizes, fits a*n^p to each, extrapolates to n_target.  `known`: {n: seconds} already measured."""
    known = dict(known or {})
    pts = []
    sizes = [nn for nn in REF_FIT_ZQ if nn <= n_target] or [n_target]   # a workload below the smallest fit size is its own sample
    for nn in sizes:
        M = O.gen_testcc(nn)[2] if nn == 500 else O.gen_sym(nn, 32)      # n = 500: the test.cc matrix itself (config 1)
        if nn in known:
            tz = known[nn]
        else:
            t0 = time.perf_counter()
            ref.zquatev(M)
            tz = time.perf_counter() - t0
        th = None
        if nn in REF_FIT_ZHEEV:
            t0 = time.perf_counter()
            ref.zheev(M)
            th = time.perf_counter() - t0
        pts.append({"n": nn, "n2": 2 * nn, "zquatev_s": tz, "zheev_2n_s": th})
    az, pz = fit_power([p["n"] for p in pts], [p["zquatev_s"] for p in pts])
    hp = [p for p in pts if p["zheev_2n_s"] is not None]
    ah, ph = fit_power([p["n"] for p in hp], [p["zheev_2n_s"] for p in hp]) if hp else (None, None)
    return {"points": pts, "fit_zquatev": {"a": az, "p": pz}, "fit_zheev": {"a": ah, "p": ph},
            "zquatev_extrapolated_s": az * n_target ** pz, "zheev_extrapolated_s": (ah * n_target ** ph) if ah else None}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n2", type=int, default=int(os.environ.get("ZQ_BENCH_N2", "32768")))
    ap.add_argument("--nb", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        load = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def traffic_model(n, world):
    """DRAM traffic of the K1 launches of one step, from the committed `ncu --set full` samples
    (profiles/r02_k1_traffic.jsonl: one line per captured launch with its trailing size m, dram bytes read + written
    and algorithmic bytes): the measured traffic/algorithmic ratio is interpolated in m and integrated over the
    n - 1 launches of the step.  Returns (traffic_bytes_per_step, info) or (None, why)."""
    p = os.path.join(ROOT, "profiles", "r02_k1_traffic.jsonl")
    try:
        rows = sorted((json.loads(l) for l in open(p) if l.strip()), key=lambda r: r["m"])
        ms_ = [r["m"] for r in rows]
        ratios = [(r["dram_read_bytes"] + r["dram_write_bytes"]) / r["algorithmic_bytes"] for r in rows]
    except Exception as ex:
        return None, "no ncu samples (%s)" % str(ex)[:80]
    import bisect

    def ratio(m):
        i = bisect.bisect_left(ms_, m)
        if i == 0:
            return ratios[0]
        if i >= len(ms_):
            return ratios[-1]
        f = (m - ms_[i - 1]) / (ms_[i] - ms_[i - 1])
        return ratios[i - 1] + f * (ratios[i] - ratios[i - 1])

    tot = 0.0
    for k in range(n - 1):
        m = n - k - 1
        tot += (16.0 * m * m / world + 64.0 * m) * ratio(m)
    return tot, {"source": "profiles/r02_k1_traffic.jsonl (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                 "sampled_m": ms_, "traffic_over_algorithmic": ratios}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: rank 0 only; the unmodified reference on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import zquatev_oracle as O
    n = args.n2 // 2
    if not O.RefLib.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libzquatev_ref.so not built"}))
        return
    ref = O.RefLib()
    cores = os.cpu_count() or 1
    ref.set_threads(cores)
    ns = min(REF_SAMPLE_N, n)
    M = O.gen_sym(ns, 32)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        eig, out, info = ref.zquatev(M)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    per = sum(times) / len(times)
    # once per run: the other sizes + zheev, power-law fit, extrapolation with the FITTED exponent
    fit = cpu_reference_points(ref, O, n, known={ns: per})
    p = fit["fit_zquatev"]["p"]
    val = per * (n / ns) ** p
    sample = (f"reference ts::zquatev (oracle/_ref, OpenBLAS {cores} threads) at 2n={2 * ns}: {per:.3f} s/solve measured per step; "
              f"EXTRAPOLATED to 2n={args.n2} by (n/{ns})^p with p = {p:.3f} fitted over n = "
              f"{[q['n'] for q in fit['points']]} (a cubic law would give {per * (n / ns) ** 3:.0f} s)")
    line = {"metric": METRIC % args.n2, "value": val, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": bench_config(args.n2, args.nb), "extrapolated": True,
            "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "kind": "reference", "sample": sample, "fit": fit},
            "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "tflops_canonical": 164.0 / 3.0 * n ** 3 / val * 1e-12}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def make_input(n, device, seed=32):
    """G_sym left half (A;B) in column-major memory: tensor [n cols][2n rows] complex128.
    Only the lower triangles are read by the solver, so no symmetrisation pass is needed; the
    diagonal of A is made real."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    left = torch.empty((n, 2 * n), dtype=torch.complex128, device=device)
    chunk = max(1, (1 << 27) // (2 * n))
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        X = torch.rand((c1 - c0, 2 * n, 2), dtype=torch.float64, device=device, generator=g) - 0.5
        left[c0:c1] = torch.view_as_complex(X)
    idx = torch.arange(n, device=device)
    left[idx, idx] = left[idx, idx].real.to(torch.complex128)
    return left


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    import zquatev_b200 as z

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
        from zquatev_b200 import dist as zd
        zd.init_from_torch()                                      # the solver's own NCCL communicator
    n2 = args.n2
    n = n2 // 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    left0 = make_input(n, dev)                                    # pristine input, stays in HBM
    work = torch.empty((n2, n2), dtype=torch.complex128, device=dev)   # [col][row] = column-major 2n x 2n
    eig = torch.zeros(n, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        work[:n].copy_(left0)                                     # restore the (destroyed) input: D2D, 16 n^2 B
        # N > 1: ONE collective call; reduction 1-D block-cyclic over the ranks, back-transformation
        # split by eigenvector columns, result complete on every rank (NCCL inside the library)
        info = z.zquatev_device(n2, work.data_ptr(), n2, eig.data_ptr(), nb=args.nb, stream=stream, sync=True,
                                dist=world > 1)
        return info

    for _ in range(args.warmup):
        info = step_device()
        assert info == 0, info
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    launches = 0
    for _ in range(args.steps):
        info = step_device()
        launches += int(z.last_phases()["launches"]) + 1
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    phases = z.last_phases()
    if world > 1:
        phases["gather"] = z.last_gather_ms()                    # NCCL gather of the eigenvector shards (inside "backtransform")

    # ---- parity of the LAST TIMED result at the full size, outside the timed region (checker: torch/cuBLAS) ----
    # the checks the reference's own test prints (test.cc:104-112) in the north_star normalisation
    from tests import gpu_util as GU
    DE = GU.build_DE(left0)

    def allsum(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    quality = GU.device_quality(left0, work, eig, col_chunk=2048, rank=rank, world=world, reduce=allsum, DE=DE)
    trace_err = quality["trace_err"]
    normA = quality["eig_absmax"]                                 # ||A||_2 of the Hermitian matrix
    if world > 1:
        lst = [torch.zeros_like(eig) for _ in range(world)]
        dist.all_gather(lst, eig)
        quality["eig_identical_across_ranks"] = all(torch.equal(lst[0], t) for t in lst)
        # run-to-run: a second collective solve must reproduce the first bit for bit (fixed-order sums, no atomics)
        eig_first = eig.clone()
        csum_first = work.view(torch.float64).sum().item()
        step_device()
        quality["bitwise_reproducible"] = bool(torch.equal(eig_first, eig)) and work.view(torch.float64).sum().item() == csum_first
        # the same matrix through the single-GPU path on this rank (no collective): eigenvalues must agree
        eig1 = torch.zeros_like(eig)
        work[:n].copy_(left0)
        info1 = z.zquatev_device(n2, work.data_ptr(), n2, eig1.data_ptr(), nb=args.nb, stream=stream, sync=True, dist=False)
        quality["eig_vs_single_gpu_rel"] = (eig1 - eig_first).abs().max().item() / normA if info1 == 0 else None
        barrier()
        assert quality["eig_identical_across_ranks"], "ranks disagree on the eigenvalues"
        assert quality["bitwise_reproducible"], "collective solve is not bit-reproducible run to run"
        assert quality["eig_vs_single_gpu_rel"] is not None and quality["eig_vs_single_gpu_rel"] <= 1e-12, quality
    quality["thresholds"] = "asserted: residual < 0.5, orthogonality < 1.5, pairing == 0, ascending, sumsq_relerr < 1e-12"
    assert quality["pairing"] == 0.0 and quality["residual"] < 0.5 and quality["orthogonality"] < 1.5, quality
    assert quality["ascending"] and quality["sumsq_relerr"] < 1e-12 and trace_err <= 1e-13 * n * normA, quality

    # ---- roofline of the dominant kernel (K1): one extra profiled step, per-launch CUDA events ----
    roof = None
    z.set_profiling(True)                                         # collective for N > 1: every rank runs it
    step_device()
    torch.cuda.synchronize()
    ph = z.last_phases()
    k4_ms = z.last_trailing_ms()                                  # event pair around every trailing-update GEMM (any N)
    z.set_profiling(False)
    barrier()
    if rank == 0:
        k1_ms = ph["k1_matvec"]
        # lower triangles of D,E + v,y; with N ranks each rank streams 1/N of the column blocks
        alg_bytes = sum(16.0 * (n - k - 1) ** 2 / world + 64.0 * (n - k - 1) for k in range(n - 1))
        peaks, src = measured_peaks()
        ach = alg_bytes / (k1_ms * 1e-3) * 1e-9 if k1_ms > 0 else None
        traffic, traffic_info = traffic_model(n, world)
        roof = {"kernel": "k_matvec (K1 quaternion-Hermitian mat-vec, lower triangles)", "bound": "hbm",
                "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (ach / peaks["hbm_gbs"]) if ach else None,
                "traffic": traffic, "traffic_info": traffic_info,
                "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None, "peak_source": src, "launches": n - 1, "k1_ms_per_step": k1_ms,
                "share_of_step": k1_ms / ph["device_total"] if ph["device_total"] else None,
                "algorithmic_bytes_per_step": alg_bytes,
                "note": "algorithmic bytes = 16 m^2 per column (lower triangles only; SURVEY 8d's full-storage figure is 32 m^2)"}

    # ---- e2e through the C ABI with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        try:
            import psutil
            need = world * (16.0 * n2 * n2 + 16.0 * n2 * n) * 1.15            # every rank pins the full array
            if psutil.virtual_memory().available < need:
                raise MemoryError(f"host RAM: need {need / 2**30:.0f} GiB for {world} pinned arrays")
            host = torch.empty((n2, n2), dtype=torch.complex128, pin_memory=True)
            host0 = left0.cpu()
            eig_h = np.zeros(n2)
            opt_times = []
            del work
            torch.cuda.empty_cache()
            reps = max(1, min(args.steps, 2))
            for it in range(1 + reps):
                host[:n].copy_(host0)
                barrier()
                t0 = time.perf_counter()
                opt = z.ZqOptions(1, 0, args.nb, None, 1, 0, 0, 1 if world > 1 else 0, 1)
                info = z.lib().zquatev_b200_ex(n2, ctypes.c_void_p(host.data_ptr()), n2, eig_h.ctypes.data, ctypes.byref(opt))
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                assert info == 0, info
                if it >= 1:
                    opt_times.append(dt)
            # the host result of the last e2e call: eigenvalues against the device-resident solve, and the residual of a
            # spread sample of host columns (left AND right half) computed on the device
            e2e_chk = {"eig_max_abs_diff_vs_device": float(np.max(np.abs(eig_h[:n] - eig.cpu().numpy())))}
            if rank == 0:
                cols = sorted(set(int(c) for c in np.linspace(0, n2 - 1, 16)))
                Xs = torch.stack([host[c] for c in cols]).to(dev).T                  # 2n x 16
                lam_s = torch.tensor([eig_h[c % n] for c in cols], dtype=torch.float64, device=dev)
                Xa, Xb = Xs[:n], Xs[n:]
                top = DE[0] @ Xa - (DE[1] @ Xb.conj()).conj()
                bot = DE[1] @ Xa + (DE[0] @ Xb.conj()).conj()
                rs = torch.sqrt(torch.linalg.norm(top - Xa * lam_s[None, :]) ** 2 + torch.linalg.norm(bot - Xb * lam_s[None, :]) ** 2).item()
                e2e_chk["sample_cols"] = len(cols)
                e2e_chk["sample_residual"] = rs / (len(cols) ** 0.5 * quality["fro_norm"] * 2.220446049250313e-16 * (n2 ** 0.5))
                assert e2e_chk["sample_residual"] < 0.5 and e2e_chk["eig_max_abs_diff_vs_device"] <= 1e-12 * normA, e2e_chk
            te = torch.tensor([sum(opt_times) / len(opt_times)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            # H2D: only the lower triangles of A and B travel, in 256-column blocks (upload_lower, csrc/solver.cu)
            # (N > 1: the same bytes in total, 1/N of them through each rank's host link)
            h2d = sum(2 * (n - c0) * min(256, n - c0) * 16 for c0 in range(0, n, 256)) if n >= 1024 else 16 * n2 * n
            e2e = {"value": te.item(), "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16 * n2 * n2 + 8 * n,
                   "phases_ms": z.last_phases(), "check": e2e_chk,
                   "note": "host-pointer C ABI call zquatev_b200_ex (== ts::zquatev), pinned host array; H2D = lower triangles of (A; B) only; "
                           "eigenvector column blocks are downloaded on a copy stream while the next block is back-transformed" + ("; N > 1: every rank uploads 1/N of the lower triangles (equal triangle areas) and NVLink carries the rest; each rank back-transforms its column shard in sub-blocks whose exchange (ncclSend/Recv to rank 0) and download overlap the next sub-block; rank 0 downloads all 2n columns, the others their own column blocks (host_result = 1)" if world > 1 else "")}
        except Exception as ex:   # e.g. not enough pinned host memory on the box
            e2e = {"value": None, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference ts::zquatev AND zheev(2n) at BASELINE.md 3's sizes ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from oracle import zquatev_oracle as O
            if O.RefLib.available():
                ref = O.RefLib()
                cores = os.cpu_count() or 1
                ref.set_threads(cores)
                fit = cpu_reference_points(ref, O, n)
                val = fit["zquatev_extrapolated_s"]
                cpu = {"value": val, "unit": "s", "cores": cores, "kind": "reference", "extrapolated": True, "fit": fit,
                       "sample": "reference ts::zquatev (oracle/_ref) measured on %d host threads at n = %s: %s s; zheev(2n) at n = %s: %s s; "
                                 "EXTRAPOLATED to 2n=%d with the fitted law t = a*n^p, p = %.3f (zheev: p = %.3f -> %.0f s)" % (
                                     cores, [q["n"] for q in fit["points"]], ["%.2f" % q["zquatev_s"] for q in fit["points"]],
                                     [q["n"] for q in fit["points"] if q["zheev_2n_s"] is not None],
                                     ["%.2f" % q["zheev_2n_s"] for q in fit["points"] if q["zheev_2n_s"] is not None],
                                     n2, fit["fit_zquatev"]["p"], fit["fit_zheev"]["p"] or 0.0, fit["zheev_extrapolated_s"] or 0.0)}
        except Exception as ex:
            cpu = {"value": None, "unit": "s", "cores": 0, "kind": "reference", "sample": "failed: " + str(ex)[:120]}

    if rank == 0:
        sec = ms_step * 1e-3
        line = {"metric": METRIC % n2, "value": sec, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": bench_config(n2, args.nb),
                "setup": {"parallelism": "1 GPU" if world == 1 else f"{world} GPUs: reduction 1-D block-cyclic (64-column blocks), per-column reflector broadcast + partial mat-vec all-reduce by " + ("peer-memory stores fused into the panel kernels (CUDA IPC over NVLink)" if z.lib().zquatev_b200_dist_transport() == 2 else "NCCL collectives") + ", D&C replicated below the top merge, back-transform sharded by eigenvector columns, NCCL gather of the result"},
                "tflops_canonical": 164.0 / 3.0 * n ** 3 / sec * 1e-12,
                "phases_ms": phases, "trace_error": trace_err, "quality": quality, "gpu_launches": launches, "clocks": clocks,
                "roofline": roof,
                # second roofline (FP64 tensor path): executed GEMM flops of the back-transformation (32 n^3 / N per rank)
                # over its CUDA-event time; peak = DMMA rate measured on this pool with tools/fp64_peak.cu
                "roofline_fp64": {"kernel": "K6 back-transformation: " + GEMM_NAME + ", DMMA m8n8k4",
                                  "bound": "tensor",
                                  # EXECUTED DMMA flops: GEMM_EXEC x the canonical 32 n^3 (see GEMM_EXEC above)
                                  "executed_over_canonical": GEMM_EXEC,
                                  "achieved": GEMM_EXEC * 32.0 * n ** 3 / world / (phases["backtransform"] * 1e-3) * 1e-12 if phases["backtransform"] else None,
                                  "peak": 37.1, "unit": "TFLOP/s",
                                  "frac": GEMM_EXEC * 32.0 * n ** 3 / world / (phases["backtransform"] * 1e-3) * 1e-12 / 37.1 if phases["backtransform"] else None,
                                  "canonical_tflops": 32.0 * n ** 3 / world / (phases["backtransform"] * 1e-3) * 1e-12 if phases["backtransform"] else None,
                                  "peak_source": "own measurement (profiles/r01_fp64_peak.jsonl: DMMA 37.1, DFMA 36.9 TFLOP/s); MEASURED_PEAKS.json has no FP64 entry",
                                  "note": "achieved / frac = EXECUTED tensor-pipe flops over the phase time (operand staging, T factors, pairing and, for N > 1, the NCCL gather included); canonical_tflops = useful work rate (32 n^3 / t), which exceeds the machine peak because the 8-product scheme needs half the real multiplications; the in-loop component sums (DADD) run on the same physical pipe and are not counted (ncu: profiles/r02_ncu_qgemm.md)"},
                "cpu_baseline": cpu, "e2e": e2e}
        if k4_ms > 0:
            # third roofline: the trailing rank-2k update [D;E] -= L R^H (K4), event pair around each of its n/nb launches in
            # the profiled step.  canonical flops = 32 m^2 kb per panel (lower triangles of D and E, K = 4 kb complex)
            nbb = args.nb or 64
            f4 = 0.0
            for j0 in range(0, n - 1, nbb):
                kb = min(nbb, n - 1 - j0)
                m = n - (j0 + kb)
                if m > 0:
                    f4 += 32.0 * m * m * kb
            ex = GEMM_EXEC if n >= 1024 else 1.0
            line["roofline_fp64_trailing"] = {
                "kernel": "K4 trailing rank-2k update, lower: " + (GEMM_NAME if n >= 1024 else "k_zgemm_mma (4 products)") + ", DMMA m8n8k4", "bound": "tensor",
                "executed_over_canonical": ex,
                "achieved": ex * f4 / world / (k4_ms * 1e-3) * 1e-12, "peak": 37.1, "unit": "TFLOP/s",
                "frac": ex * f4 / world / (k4_ms * 1e-3) * 1e-12 / 37.1, "canonical_tflops": f4 / world / (k4_ms * 1e-3) * 1e-12,
                "note": "per rank: each rank updates the 64-column blocks it owns (1/N of the flops)",
                "k4_ms_per_step": k4_ms, "share_of_step": k4_ms / ph["device_total"] if ph["device_total"] else None,
                "launches": (n - 1 + nbb - 1) // nbb,
                "peak_source": "own measurement (profiles/r01_fp64_peak.jsonl)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        from zquatev_b200 import dist as zd
        zd.finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
