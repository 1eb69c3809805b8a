#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 robust-PCA hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c4|c2|c3]

One "step" = one full `rpca` solve to the reference's tolerance (sqrt(eps)) on the workload:
    c4 (default, the configuration the metric is quoted on): 1 000 000 x 256 FP64, rank 10 + 5 % sparse,
        nonnegA=true; rows sharded over the N GPUs (strong scaling: the global matrix is fixed)
    c2: 100 000 x 512 FP64, rank 10 + 5 % sparse (single GPU)
    c3: rpca_ga(X, 10) on 2 000 000 x 256 (rows sharded); metric = Grassmann iterations/s
Metric: ALM iterations per second (whole job).  `value` is measured with the input resident in HBM, `e2e` through
the public host-buffer API (pinned host memory, H2D of D and D2H of A and E inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/, same LAPACK routines the Julia reference
reaches through OpenBLAS; Julia itself is not installed in this image) on a bounded row sample.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram read+write per launch of alm_fused_kernel from `ncu --set full` (profiles/), keyed by (rows per GPU, columns)
FUSED_TRAFFIC_GB = {(1_000_000, 256): 6.32}
# same for syrk_tma_kernel (profiles/r01_ncu_full_final.md: 3.57 GB read + 0.01 GB written for S = 2.05 GB)
SYRK_TRAFFIC_GB = {(1_000_000, 256): 3.58}
# and for the two streaming-epilogue kernels together (profiles/r01_ncu_full_stream_split.md: 2.14 + 8.27 GB at RP = 12)
STREAM_TRAFFIC_GB = {(1_000_000, 256): 10.41}

FP64_TENSOR_PEAK_TFLOPS = 37.1      # measured on this pool's B200 (tools/microbench.cu -> profiles/r01_microbench_fp64_hbm.log)

WORKLOADS = {
    "c4": dict(kind="rpca", M=1_000_000, N=256, rank=10, frac=0.05, seed=4, nonneg=True,
               name="rpca 1Mx256 FP64 rank10+5%sparse nonnegA=true (BASELINE configs[3])"),
    "c2": dict(kind="rpca", M=100_000, N=512, rank=10, frac=0.05, seed=2, nonneg=False,
               name="rpca 100kx512 FP64 rank10+5%sparse (BASELINE configs[1])"),
    "c3": dict(kind="ga", M=2_000_000, N=256, rank=10, seed=3,
               name="rpca_ga(X,10) 2Mx256 FP64 10% gross outliers (BASELINE configs[2])"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ----------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on a bounded row sample
# ----------------------------------------------------------------------------------------------------------
def cpu_rpca_sample(w, sample_rows: int, iters: int):
    """iterations/s of the oracle on the first `sample_rows` rows, scaled to the full row count (cost is linear in
    the number of rows: dgesdd on M x n, n fixed, and the element-wise sweeps)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import warnings

    import numpy as np
    import tls_oracle as O
    import tlsq_b200 as T
    D = T.synth.lowrank_sparse_np(sample_rows, w["N"], w["rank"], w["frac"], w["seed"], w["nonneg"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        res = O.rpca(D, iters=iters, tol=0.0, nonnegA=w["nonneg"], lam=1.0 / math.sqrt(w["M"]))
        dt = time.perf_counter() - t0
    its_per_s_sample = res.iters / dt
    return its_per_s_sample * sample_rows / w["M"], dt, res.iters


def cpu_ga_sample(w, sample_rows: int, r: int):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tls_oracle as O
    import tlsq_b200 as T
    X, q0 = T.synth.ga_data_np(sample_rows, w["N"], w["rank"], w["seed"])
    t0 = time.perf_counter()
    _, its = O.rpca_ga(X, r, q0=q0[:, :r], exact_order=False, return_iters=True)
    dt = time.perf_counter() - t0
    return sum(its) / dt * sample_rows / w["M"], dt, sum(its)


def cpu_cores():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads") for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(args, w, rank):
    if rank != 0:
        return
    import numpy  # noqa: F401  (loads OpenBLAS so that cpu_cores() sees it)
    import scipy.linalg  # noqa: F401
    vals, dts = [], []
    if w["kind"] == "rpca":
        sample_rows, its = 50_000 if w["N"] <= 256 else 20_000, 3
        for i in range(args.warmup + args.steps):
            v, dt, _ = cpu_rpca_sample(w, sample_rows, its)
            if i >= args.warmup:
                vals.append(v); dts.append(dt)
        sample = (f"{its} ALM iterations of the oracle port on the first {sample_rows} rows per step, scaled by "
                  f"{sample_rows}/{w['M']} (cost linear in rows); LAPACK dgesdd/dgemm via SciPy OpenBLAS")
        unit = "ALM iterations/s"
    else:
        sample_rows, r = 100_000, 2
        for i in range(args.warmup + args.steps):
            v, dt, _ = cpu_ga_sample(w, sample_rows, r)
            if i >= args.warmup:
                vals.append(v); dts.append(dt)
        sample = (f"rpca_ga(X,{r}) of the oracle port on the first {sample_rows} rows per step, scaled by "
                  f"{sample_rows}/{w['M']}")
        unit = "GA iterations/s"
    value = sum(vals) / len(vals)
    line = {"metric": metric_name(w), "value": value, "unit": unit, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(dts) / len(dts),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"]},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cpu_cores(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    EMIT(json.dumps(line))


def metric_name(w):
    return "rpca_alm_iterations_per_s" if w["kind"] == "rpca" else "rpca_ga_iterations_per_s"


# ----------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------
def run_b200(args, w, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    import tlsq_b200 as T

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        T.init_distributed(local_rank)
    M, N = w["M"], w["N"]
    r0, r1 = T.synth.shard_rows(M, world, rank, align=T.synth.CHUNK_ROWS)
    m = r1 - r0
    lam = 1.0 / math.sqrt(max(M, N))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if w["kind"] == "rpca":
        D = T.synth.lowrank_sparse_cuda(r0, r1, N, dev, w["rank"], w["frac"], w["seed"], w["nonneg"])
        kw = dict(nonnegA=w["nonneg"], lam=lam)

        def step():
            A, E, s, sv, info = T.rpca(D, return_info=True, **kw)
            return info["iters"]
    else:
        X, q0 = T.synth.ga_data_cuda(r0, r1, N, dev, w["rank"], w["seed"])

        def step():
            Q, info = T.rpca_ga(X, w["rank"], q0=q0, return_info=True)
            return sum(info["iters"])

    # ---- device-resident timing ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = T.launch_count(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters_total = 0
    for _ in range(args.steps):
        iters_total += step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = T.launch_count(local_rank) - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = iters_total / (ms * 1e-3)

    # ---- one profiled solve: per-phase device times for the roofline ------------------------------------------
    T.set_profiling(True, local_rank)
    it_prof = step()
    prof = T.get_profile(local_rank)
    T.set_profiling(False, local_rank)

    # ---- end-to-end through the host-buffer API (pinned host memory) ----------------------------------------
    e2e = None
    if w["kind"] == "rpca":
        Dh = torch.empty((N, m), dtype=torch.float64, pin_memory=True)
        Dh.copy_(D.t())
        Dn = Dh.numpy().T                                          # column-major (m, N) view of pinned memory
        Ah = torch.empty((N, m), dtype=torch.float64, pin_memory=True).numpy().T
        Eh = torch.empty((N, m), dtype=torch.float64, pin_memory=True).numpy().T
        del D
        torch.cuda.empty_cache()

        def step_e2e():
            _, _, _, _, info = T.rpca(Dn, return_info=True, want_svd=False, out=(Ah, Eh), **kw)
            return info["iters"]
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        it_e = 0
        nst = max(1, min(args.steps, 3))
        for _ in range(nst):
            it_e += step_e2e()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": it_e / float(dt.item()), "unit": "ALM iterations/s", "h2d_bytes_per_step": int(m * N * 8),
               "d2h_bytes_per_step": int(2 * m * N * 8), "ms_per_step": 1e3 * float(dt.item()) / nst,
               "api": "tlsq_b200.rpca(numpy pinned) -> tlsq_rpca_f64 (A and E returned to the host)"}
    else:
        Xh = torch.empty((N, m), dtype=torch.float64, pin_memory=True)
        Xh.copy_(X.t())
        Xn, q0n = Xh.numpy().T, np.asfortranarray(q0.cpu().numpy())

        def step_e2e():
            _, info = T.rpca_ga(Xn, w["rank"], q0=q0n, return_info=True)
            return sum(info["iters"])
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        it_e = step_e2e()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": it_e / float(dt.item()), "unit": "GA iterations/s", "h2d_bytes_per_step": int(m * N * 8 + m * w["rank"] * 8),
               "d2h_bytes_per_step": int(m * w["rank"] * 8), "ms_per_step": 1e3 * float(dt.item())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    line = {"metric": metric_name(w), "value": value,
            "unit": "ALM iterations/s" if w["kind"] == "rpca" else "GA iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "rows_per_gpu": m, "cols": N,
                       "l2": "inputs (>=0.4 GB per pass) larger than the 126 MB L2"},
            "iters_per_step": iters_total / args.steps, "time_to_converge_s": ms * 1e-3 / args.steps,
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e}
    if w["kind"] == "rpca":
        g_ms, g_n = prof["gram"]
        e_ms, e_n = prof["epilogue"]
        j_ms, j_n = prof["eig"]
        f_ms, f_n = prof.get("fused", (0.0, 0))
        svp = w["rank"]
        S = m * N * 8
        gram_flops = m * N * (N + 1)                               # SYRK count per launch
        epi_flops = 4 * m * N * svp + 12 * m * N                   # low-rank reconstruction + element-wise (8d)
        peak_src64 = ("measured FP64 DMMA peak, profiles/r01_microbench_fp64_hbm.log "
                      "(MEASURED_PEAKS.json has no FP64 figure)")
        if f_n > 0:
            # one-pass pipeline: the dominant kernel is alm_fused_kernel (epilogue of iteration k + Gram of iteration
            # k+1).  Both halves run on the FP64 pipe (DMMA and DFMA share it), so the bound is FP64 arithmetic.
            t_f = f_ms / f_n * 1e-3
            ach = (gram_flops + epi_flops) / t_f * 1e-12
            line["roofline"] = {"kernel": "alm_fused_kernel (epilogue of iteration k + DMMA Gram of iteration k+1, "
                                          "one HBM pass)", "bound": "tensor", "achieved": ach,
                                "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_TENSOR_PEAK_TFLOPS,
                                "peak_source": peak_src64, "algorithmic_flops": "M*N*(N+1) + 4*M*N*svp + 12*M*N "
                                "(SURVEY.md 8d: F_gram + F_epi)", "traffic": FUSED_TRAFFIC_GB.get((m, N)),
                                "traffic_unit": "GB per launch (dram read+write, ncu --set full, profiles/)",
                                "avg_launch_ms": t_f * 1e3, "launches_profiled": f_n,
                                "hbm_view": {"algorithmic_GBps": 9 * S / t_f * 1e-9, "moved_GBps": 3 * S / t_f * 1e-9,
                                             "peak": hbm, "note": "9S = SURVEY 8d two-pass count; the kernel moves 3S "
                                                                   "(reads D, Y; writes Y)"}}
        else:
            t_gram = g_ms / max(g_n, 1) * 1e-3
            ach = gram_flops / t_gram * 1e-12 if t_gram > 0 else 0.0
            line["roofline"] = {"kernel": "syrk_tma_kernel (DMMA SYRK of the SVT input)", "bound": "tensor",
                                "achieved": ach, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                                "frac": ach / FP64_TENSOR_PEAK_TFLOPS, "peak_source": peak_src64,
                                "algorithmic_flops": "M*N*(N+1) (SYRK count, SURVEY.md 8d)",
                                "traffic": SYRK_TRAFFIC_GB.get((m, N)),
                                "traffic_unit": "GB per launch (dram read+write, ncu --set full, profiles/)",
                                "avg_launch_ms": t_gram * 1e3}
            t_epi = e_ms / max(e_n, 1) * 1e-3
            line["roofline_epilogue"] = {"kernel": "alm_stream_kernel<PH=1> + <PH=2> (T projection, element-wise pass)",
                                         "bound": "hbm", "traffic": STREAM_TRAFFIC_GB.get((m, N)),
                                         "achieved": 6 * S / t_epi * 1e-9 if t_epi > 0 else 0.0, "peak": hbm,
                                         "unit": "GB/s", "frac": (6 * S / t_epi * 1e-9) / hbm if t_epi > 0 else 0.0,
                                         "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                                         "algorithmic_bytes": "6*S (reads D,A,Y; writes A,E,Y; SURVEY.md 8d)",
                                         "avg_launch_ms": t_epi * 1e3}
        # per-iteration roofline:  max(F_gram/P64, 3S/BW) + max(F_epi/P64, 6S/BW)
        t_roof = max(gram_flops / (FP64_TENSOR_PEAK_TFLOPS * 1e12), 3 * S / (hbm * 1e9)) + \
            max(epi_flops / (FP64_TENSOR_PEAK_TFLOPS * 1e12), 6 * S / (hbm * 1e9))
        t_iter = ms * 1e-3 / max(iters_total, 1)
        line["iteration_roofline"] = {"roofline_ms": t_roof * 1e3, "measured_ms": t_iter * 1e3,
                                      "frac": t_roof / t_iter if t_iter > 0 else 0.0,
                                      "phase_ms_per_iter": {k: v[0] / max(it_prof, 1) for k, v in prof.items()},
                                      "eig_avg_ms": j_ms / max(j_n, 1)}
    else:
        s_ms, s_n = prof["ga_sweep"]
        t_sw = s_ms / max(s_n, 1) * 1e-3
        bytes_ = m * N * 8
        line["roofline"] = {"kernel": "ga_sweep_kernel<GA_PASS>", "bound": "hbm",
                            "achieved": bytes_ / t_sw * 1e-9 if t_sw > 0 else 0.0, "peak": hbm, "unit": "GB/s",
                            "frac": (bytes_ / t_sw * 1e-9) / hbm if t_sw > 0 else 0.0,
                            "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})", "traffic": None,
                            "avg_launch_ms": t_sw * 1e3}
    # CPU baseline beside it (rank 0, N == 1 only), bounded sample
    if world == 1 and not args.no_cpu:
        if w["kind"] == "rpca":
            sample_rows = 50_000 if N <= 256 else 20_000
            v, dt, its = cpu_rpca_sample(w, sample_rows, 3)
            sample = (f"{its} ALM iterations of the oracle port (LAPACK dgesdd/dgemm via SciPy OpenBLAS) on the first "
                      f"{sample_rows} rows ({dt:.1f} s), scaled by {sample_rows}/{M}")
            unit = "ALM iterations/s"
        else:
            v, dt, its = cpu_ga_sample(w, 100_000, 2)
            sample = f"rpca_ga(X,2) of the oracle port on the first 100000 rows ({dt:.1f} s), scaled by 100000/{M}"
            unit = "GA iterations/s"
        line["cpu_baseline"] = {"value": v, "unit": unit, "cores": cpu_cores(), "kind": "port", "sample": sample}
    EMIT(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _quiet_stdout():
    """Route fd 1 to stderr while the benchmark runs (NCCL prints its version banner on stdout when NCCL_DEBUG is set)
    and hand back a writer for the one JSON line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(text: str):
        sys.stdout.flush()
        os.write(real, (text + "\n").encode())
    return emit


EMIT = None


def main():
    global EMIT
    EMIT = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override the row count (debugging)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)
    w = dict(WORKLOADS[args.workload])
    if args.rows:
        w["M"] = args.rows
        w["name"] += f" [rows overridden to {args.rows}]"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, w, rank)
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    run_b200(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
