#!/usr/bin/env python
"""bench.py -- KKT factor+solves/sec on the BASELINE.json headline configuration.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one interior-point iteration's KKT work on the n=16384, m=262144 dense QP (C4):
NT scaling of (v,s) -> Atil = F^-T A -> H = Q + Atil'Atil (DMMA SYRK) -> [NCCL all-reduce of the
partial Gram matrices when A is row-sharded over N GPUs] -> Cholesky -> k=2 solves (predictor +
corrector right-hand sides), each solve including its two A mat-vecs.  `value` times that with all
inputs resident in HBM; `e2e` times the same call sequence through the public API with HOST
buffers (h2d of v,s and the right-hand sides, d2h of lambda and the solutions inside the timed
region).  The row shards make N>1 a strong-scaling run of the same global problem.

`--impl reference` times the CPU restatement of the reference path (oracle/, NumPy + OpenBLAS
LAPACK: the same BLAS/LAPACK routines Julia's LinearAlgebra calls) on the host cores; Julia is
not installed in this image so the reference itself cannot run.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kkt_factor_plus_solves_per_sec"
UNIT = "kkt_units/s"
CONFIGS = {"C4": (16384, 262144), "C2": (8192, 16384), "C1": (1000, 1000)}
NSOLVES = 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4", choices=list(CONFIGS))
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--m", type=int, default=0)
    ap.add_argument("--no-solve", action="store_true", help="skip the full time-to-1e-8 solve")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--py-driver", action="store_true", help="also time the Python host driver (conicIP) on the solve")
    return ap.parse_args()


def workload_name(cfg, n, m):
    return f"{cfg}: dense QP n={n}, m={m} inequality rows, K=R^m, Q=diag, k={NSOLVES} solves/unit"


# ------------------------------------------------------------------------------- CPU arm
def cpu_unit_time(n, m, budget_rows=None, chol_n=None, seed=0):
    """Time one KKT unit of the oracle port on the host cores on a bounded sample and
    extrapolate: SYRK is linear in m (exact in flops), Cholesky cubic in n.
    Returns (unit_seconds, description, threads)."""
    import numpy as np
    import scipy.linalg as sla
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        threads = os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    ms = min(m, budget_rows or max(256, int(1.5e12 / (2.0 * n * n))))     # ~1.5e12 flop of dgemm
    ns = min(n, chol_n or 8192)
    A = rng.standard_normal((ms, n)) / math.sqrt(n)
    f = rng.uniform(0.5, 2.0, ms)
    t0 = time.perf_counter()
    Atil = A / f[:, None]                       # F^-T A            (src/kktsolvers.jl:33)
    H = Atil.T @ Atil                           # Atil'Atil         (:34)
    H[np.diag_indices(n)] += 1.5
    t_syrk = time.perf_counter() - t0
    Hs = H[:ns, :ns] + ms * np.eye(ns)
    t0 = time.perf_counter()
    L = sla.cholesky(Hs, lower=True, check_finite=False, overwrite_a=True)
    t_chol = time.perf_counter() - t0
    rv = rng.standard_normal(ms)
    ry = rng.standard_normal(ns)
    t0 = time.perf_counter()
    for _ in range(NSOLVES):                    # pivot algebra, src/kktsolvers.jl:324-332
        t1 = rv / (f * f)
        rhs = ry + (A.T @ t1)[:ns]
        y = sla.solve_triangular(L, rhs, lower=True, check_finite=False)
        dy = sla.solve_triangular(L.T, y, lower=False, check_finite=False)
        dv = t1 - (A[:, :ns] @ dy) / (f * f)
    t_solve = time.perf_counter() - t0
    del dv
    # extrapolate: gemv/syrk linear in m; cholesky cubic, triangular solves quadratic in n
    gemv_part = t_solve * 0.5
    unit = t_syrk * (m / ms) + t_chol * (n / ns) ** 3 + gemv_part * (m / ms) + (t_solve - gemv_part) * (n / ns) ** 2
    desc = (f"oracle port (NumPy/OpenBLAS dgemm Atil'Atil + LAPACK dpotrf/dtrsv), {threads} threads: SYRK on "
            f"{ms} of {m} rows x linear extrapolation ({t_syrk:.2f}s), Cholesky n={ns} x cubic extrapolation "
            f"({t_chol:.2f}s), {NSOLVES} solves ({t_solve:.2f}s)")
    return unit, desc, threads


def run_reference(args, n, m):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must use all the host threads it can
    ncpu = str(os.cpu_count() or 1)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = ncpu
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=int(ncpu))
    except Exception:
        pass
    times = []
    desc, threads = "", 1
    for i in range(args.warmup + args.steps):
        u, desc, threads = cpu_unit_time(n, m, seed=i)
        if i >= args.warmup:
            times.append(u)
    unit = sum(times) / len(times)
    val = 1.0 / unit
    line = {
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": unit * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args.config, n, m)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
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
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------- B200 arm
def run_b200(args, n, m):
    import numpy as np
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist

    import conicip_b200 as cb
    from conicip_b200 import problems as P
    from conicip_b200.dist import TorchReducer, init_engine_comm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the b200 arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    assert m % (world * 32) == 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t_setup = time.time()
    prob = P.config4_device(n=n, m=m, seed=4, rank=rank, nranks=world)
    m_loc = prob["m_loc"]
    qdiag = prob["qdiag"].cpu().numpy()
    eng = cb.Engine(sp.diags(qdiag).tocsr(), prob["At"].t(), None, prob["cone_dims"])
    b_loc, c_vec = prob["b"], prob["c"]
    del prob["At"]
    torch.cuda.empty_cache()
    eng._bind_stream()
    if world > 1:
        init_engine_comm(eng)
    t_setup = time.time() - t_setup

    g = torch.Generator(device="cuda")
    g.manual_seed(100 + rank)
    rnd = lambda k, lo=0.0: torch.rand(k, generator=g, dtype=torch.float64, device="cuda") + lo
    pts = [(rnd(m_loc, 0.5), rnd(m_loc, 0.5)) for _ in range(2)]
    g.manual_seed(7)                                  # n-vectors are replicated across ranks
    rys = [torch.randn(n, generator=g, dtype=torch.float64, device="cuda") for _ in range(NSOLVES)]
    g.manual_seed(200 + rank)
    rvs = [torch.randn(m_loc, generator=g, dtype=torch.float64, device="cuda") for _ in range(NSOLVES)]

    def step_device(i):
        v, s = pts[i % 2]
        eng.factor_from_point(v, s)
        for k in range(NSOLVES):
            eng.solve(rys[k], None, rvs[k])

    # ---- device-resident timing
    for i in range(args.warmup):
        step_device(i)
    st0 = eng.stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    syrk_ms, chol_ms, ar_ms, scale_ms = [], [], [], []
    e0.record()
    for i in range(args.steps):
        step_device(i)
        st = eng.stats()
        syrk_ms.append(st["ms_syrk"]); chol_ms.append(st["ms_chol"]); ar_ms.append(st["ms_allreduce"])
        scale_ms.append(st["ms_scale"])
    e1.record()
    barrier()
    ms_total = maxr(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    st1 = eng.stats()
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step

    # ---- end-to-end through the public API with host buffers
    pin = lambda t: t.cpu().pin_memory().numpy()
    h_pts = [(pin(v), pin(s)) for v, s in pts]
    h_rys, h_rvs = [pin(x) for x in rys], [pin(x) for x in rvs]
    h2d = 8 * (2 * m_loc + NSOLVES * (n + m_loc))
    d2h = 8 * (m_loc + NSOLVES * (n + m_loc))

    def step_host(i):
        v, s = h_pts[i % 2]
        lam = eng.factor_from_point(v, s)
        outs = [eng.solve(h_rys[k], None, h_rvs[k]) for k in range(NSOLVES)]
        return lam, outs

    e2e_steps = max(2, min(args.steps, 3))
    step_host(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step_host(i)
    barrier()
    e2e_ms = maxr((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e = {"value": 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(sumr(float(h2d))), "d2h_bytes_per_step": int(sumr(float(d2h))),
           "timer": "host perf_counter around the API calls (each call blocks until outputs are on the host), max over ranks"}

    # ---- roofline of the dominant kernel (gemm_nt as the SYRK)
    syrk_flops = float(m_loc) * n * n                       # algorithmic m n^2, this rank's rows
    syrk_avg_ms = sum(syrk_ms) / len(syrk_ms)
    achieved = syrk_flops / (syrk_avg_ms * 1e-3) / 1e12
    peaks = cb.measure_fp64_peaks()                 # short register-only DMMA / DFMA loops (burst clocks)
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    bmat = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ bmat
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(3):
        a @ bmat
    g1.record()
    torch.cuda.synchronize()
    dgemm_tf = 3 * 2 * 8192 ** 3 / (g0.elapsed_time(g1) * 1e-3) / 1e12
    del a, bmat
    roofline = {
        "bound": "tensor", "kernel": "gemm_nt_kernel (SYRK H = Q + Atil'Atil)", "achieved": achieved,
        "peak": dgemm_tf, "unit": "TFLOP/s", "frac": achieved / dgemm_tf, "traffic": TRAFFIC_BYTES.get((n, m_loc)),
        "peak_source": "cuBLAS DGEMM 8192^3 FP64 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
        "dmma_register_peak_tflops": peaks["dmma_tflops"], "dfma_register_peak_tflops": peaks["dfma_tflops"],
        "frac_of_dmma_register_peak": achieved / peaks["dmma_tflops"],
        "flops_per_launch": syrk_flops, "ms_per_launch": syrk_avg_ms,
        "step_breakdown_ms": {"scale_panel": sum(scale_ms) / len(scale_ms), "syrk": syrk_avg_ms,
                              "allreduce": sum(ar_ms) / len(ar_ms), "cholesky": sum(chol_ms) / len(chol_ms),
                              "solve_each": st1["ms_solve"]},
    }

    # ---- full interior-point solve: time-to-1e-8 (native loop: one cip_ipm_solve call per rank)
    solve_info = None
    if not args.no_solve:
        barrier()
        t0 = time.perf_counter()
        _, _, _, info = eng.ipm_solve(c_vec, b_loc, None, optTol=1e-8)
        barrier()
        t_solve = maxr(time.perf_counter() - t0)
        solve_info = {"time_to_1e-8_s": t_solve, "status": info["status"], "iterations": info["Iter"],
                      "factors": info["factors"], "solves": info["solves"], "prFeas": info["prFeas"],
                      "duFeas": info["duFeas"], "muFeas": info["muFeas"], "driver": "cip_ipm_solve (native)"}
        if args.py_driver:
            class _Shape:
                def __init__(self, *s):
                    self.shape = s

            def kk(Q, A, G, cd):
                def gen(F, Finvt=None):
                    st = eng.factor_resident() if isinstance(F, cb.DeviceBlock) else eng.factor(F)
                    assert st == 0, st
                    return lambda y, w, v: eng.solve(y, w, v)
                gen.engine = eng
                return gen

            barrier()
            t0 = time.perf_counter()
            sol = cb.conicIP(_Shape(n, n), c_vec.cpu().numpy(), _Shape(m_loc, n), b_loc.cpu().numpy(),
                             [("R", m_loc)], kktsolver=kk, optTol=1e-8,
                             reducer=TorchReducer() if world > 1 else None)
            barrier()
            solve_info["python_driver"] = {"time_to_1e-8_s": maxr(time.perf_counter() - t0), "status": sol.status,
                                           "iterations": sol.Iter, "factors": sol.factors, "solves": sol.solves}

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        unit, desc, threads = cpu_unit_time(n, m)
        cpu = {"value": 1.0 / unit, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, n, m), "rows_per_gpu": m_loc,
                       "l2": "inputs larger than L2 (A slab %.1f GB per GPU re-read every step)" % (m_loc * n * 8 / 1e9),
                       "setup_s": t_setup},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "full_solve": solve_info,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per SYRK launch from the committed ncu --set full
# capture (profiles/), keyed by (n, rows on the GPU); None where no capture exists.
TRAFFIC_BYTES = {
    (16384, 262144): 811.947289e9 + 1.196444e9,   # profiles/r01_ncu_syrk_band.md (C4, one GPU)
    (8192, 16384): 6.703572e9 + 0.270010e9,       # same file (C2)
}


def main():
    args = parse()
    n, m = CONFIGS[args.config]
    n, m = args.n or n, args.m or m
    if args.impl == "reference":
        run_reference(args, n, m)
    else:
        run_b200(args, n, m)


if __name__ == "__main__":
    main()
