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
          block-cyclic by 64-column blocks (per column one NCCL broadcast of the reflector and one
          all-reduce of the partial mat-vec), trailing update on owned blocks only, back-transformation
          sharded by eigenvector columns, result gathered on every rank; the tridiagonal D&C is
          replicated.
--impl reference : the UNMODIFIED reference ts::zquatev built in oracle/_ref, timed on the host
          cores on a bounded sample (2n = 2048) and scaled by n^3 to the workload (SURVEY.md 8d:
          a real 2n=32768 CPU run takes ~12 h).
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
GEMM_EXEC = 0.75 if os.environ.get("ZQ_GEMM_3M", "1") != "0" else 1.0   # executed / canonical GEMM flops
REF_SAMPLE_N = 1024          # 2n = 2048 reference run per step (~5-10 s on 16 cores)


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


def traffic_probe():
    """dram bytes of ONE K1 launch from the committed ncu --set full capture (the step average over 16383 launches of
    shrinking size is not capturable under ncu; `traffic` itself therefore stays null)"""
    p = os.path.join(ROOT, "profiles", "r01_k1_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


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
    scale = (n / ns) ** 3
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        eig, out, info = ref.zquatev(M)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    per = sum(times) / len(times)
    val = per * scale
    sample = (f"reference ts::zquatev (oracle/_ref, OpenBLAS {cores} threads) at 2n={2 * ns}: {per:.3f} s/solve, "
              f"scaled by (n/{ns})^3 = {scale:.0f} to 2n={args.n2}")
    line = {"metric": METRIC % args.n2, "value": val, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"2n={args.n2} quaternionic Hermitian eigendecomposition (values+vectors)",
                       "sample": sample},
            "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "kind": "reference", "sample": sample},
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

    # ---- sanity of the last result (cheap, outside the timed region): sum(eig) = trace(A) ----
    tr = torch.diagonal(left0[:, :n]).real.sum().item()
    trace_err = abs(eig.sum().item() - tr)

    # ---- roofline of the dominant kernel (K1): one extra profiled step, per-launch CUDA events ----
    roof = None
    z.set_profiling(True)                                         # collective for N > 1: every rank runs it
    step_device()
    torch.cuda.synchronize()
    ph = z.last_phases()
    k4_ms = z.last_trailing_ms() if world == 1 else 0.0
    z.set_profiling(False)
    barrier()
    if rank == 0:
        k1_ms = ph["k1_matvec"]
        # lower triangles of D,E + v,y; with N ranks each rank streams 1/N of the column blocks
        alg_bytes = sum(16.0 * (n - k - 1) ** 2 / world + 64.0 * (n - k - 1) for k in range(n - 1))
        peaks, src = measured_peaks()
        ach = alg_bytes / (k1_ms * 1e-3) * 1e-9 if k1_ms > 0 else None
        roof = {"kernel": "k_matvec (K1 quaternion-Hermitian mat-vec, lower triangles)", "bound": "hbm",
                "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (ach / peaks["hbm_gbs"]) if ach else None,
                "traffic": None, "traffic_probe": traffic_probe(), "peak_source": src, "launches": n - 1, "k1_ms_per_step": k1_ms,
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
            te = torch.tensor([sum(opt_times) / len(opt_times)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e = {"value": te.item(), "unit": "s", "h2d_bytes_per_step": 16 * n2 * n, "d2h_bytes_per_step": 16 * n2 * n2 + 8 * n,
                   "phases_ms": z.last_phases(),
                   "note": "host-pointer C ABI call zquatev_b200_ex (== ts::zquatev), pinned host array" + ("; every rank uploads its copy of the input, rank 0 downloads all 2n columns, the others their own column blocks" if world > 1 else "")}
        except Exception as ex:   # e.g. not enough pinned host memory on the box
            e2e = {"value": None, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the reference ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from oracle import zquatev_oracle as O
            if O.RefLib.available():
                ref = O.RefLib()
                cores = os.cpu_count() or 1
                ref.set_threads(cores)
                ns = min(REF_SAMPLE_N, n)
                M = O.gen_sym(ns, 32)
                t0 = time.perf_counter()
                ref.zquatev(M)
                dt = time.perf_counter() - t0
                sc = (n / ns) ** 3
                cpu = {"value": dt * sc, "unit": "s", "cores": cores, "kind": "reference",
                       "sample": f"reference ts::zquatev (oracle/_ref) at 2n={2 * ns}: {dt:.3f} s measured on {cores} host threads, scaled by (n/{ns})^3={sc:.0f}"}
        except Exception as ex:
            cpu = {"value": None, "unit": "s", "cores": 0, "kind": "reference", "sample": "failed: " + str(ex)[:120]}

    if rank == 0:
        sec = ms_step * 1e-3
        line = {"metric": METRIC % n2, "value": sec, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"2n={n2} quaternionic Hermitian eigendecomposition (values+vectors), G_sym seed 32",
                           "n2": n2, "nb": args.nb or 64, "l2": "inputs (16*n2*n B) larger than L2; fresh copy of the input every step",
                           "parallelism": "1 GPU" if world == 1 else f"{world} GPUs: reduction 1-D block-cyclic (64-column blocks), per-column reflector broadcast + partial mat-vec all-reduce by " + ("peer-memory stores fused into the panel kernels (CUDA IPC over NVLink)" if z.lib().zquatev_b200_dist_transport() == 2 else "NCCL collectives") + ", D&C replicated, back-transform sharded by eigenvector columns, NCCL gather of the result"},
                "tflops_canonical": 164.0 / 3.0 * n ** 3 / sec * 1e-12,
                "phases_ms": phases, "trace_error": trace_err, "gpu_launches": launches, "clocks": clocks,
                "roofline": roof,
                # second roofline (FP64 tensor path): executed GEMM flops of the back-transformation (32 n^3 / N per rank)
                # over its CUDA-event time; peak = DMMA rate measured on this pool with tools/fp64_peak.cu
                "roofline_fp64": {"kernel": "k_zgemm_3m (K6 back-transformation, DMMA m8n8k4, 3 real products per complex product)",
                                  "bound": "tensor",
                                  # EXECUTED flops: the 3M scheme performs 3/4 of the canonical 32 n^3 (ZQ_GEMM_3M=0: all of them)
                                  "achieved": GEMM_EXEC * 32.0 * n ** 3 / world / (phases["backtransform"] * 1e-3) * 1e-12 if phases["backtransform"] else None,
                                  "peak": 37.1, "unit": "TFLOP/s",
                                  "frac": GEMM_EXEC * 32.0 * n ** 3 / world / (phases["backtransform"] * 1e-3) * 1e-12 / 37.1 if phases["backtransform"] else None,
                                  "canonical_tflops": 32.0 * n ** 3 / world / (phases["backtransform"] * 1e-3) * 1e-12 if phases["backtransform"] else None,
                                  "peak_source": "own measurement (profiles/r01_fp64_peak.jsonl: DMMA 37.1, DFMA 36.9 TFLOP/s); MEASURED_PEAKS.json has no FP64 entry",
                                  "note": "phase time includes operand staging, T factors, pairing and (N > 1) the NCCL gather"},
                "cpu_baseline": cpu, "e2e": e2e}
        if world == 1 and k4_ms > 0:
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
                "kernel": "k_zgemm_3m<0,1> lower (K4 trailing rank-2k update, DMMA m8n8k4)", "bound": "tensor",
                "achieved": ex * f4 / (k4_ms * 1e-3) * 1e-12, "peak": 37.1, "unit": "TFLOP/s",
                "frac": ex * f4 / (k4_ms * 1e-3) * 1e-12 / 37.1, "canonical_tflops": f4 / (k4_ms * 1e-3) * 1e-12,
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
