#!/usr/bin/env python
"""bench.py -- the hot path of BASELINE.json on N B200s (or the CPU reference arm).

    python bench.py --gpus N --steps K --warmup W            # CUDA path (libbagpu.so)
    python bench.py --impl reference --steps K --warmup W     # CPU restatement of the reference

A *step* is one fused evaluation ``x -> (cx, vals)`` (cons! + jac_coord!, src/BALNLPModels.jl:115-206)
of the Venice-1778-shaped synthetic problem, observation-sharded over the ranks (no collective on this
path).  ``value`` = observations of the whole problem / max-over-ranks device time, inputs resident in
HBM; ``e2e`` = the same through the host-pointer C-ABI call (pinned host buffers, H2D of x and D2H of
cx/vals inside the timed region).  The LM leg reports full Levenberg-Marquardt iterations per second
on the same problem (``lm`` object).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "residual+Jacobian Mobs/s"
UNIT = "Mobs/s"


def algorithmic_bytes_per_obs(p) -> float:
    """2 int32 indices + pt2d 16 + cx 16 + vals 192 + each parameter read once."""
    return 232.0 + 8.0 * p.nvar / p.nobs


def measured_traffic(workload, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture
    (profiles/k_eval_traffic.json); only valid for the configuration it was captured on."""
    try:
        with open(os.path.join(ROOT, "profiles", "k_eval_traffic.json")) as f:
            t = json.load(f)
        if t["workload"] == workload and t["n_gpus"] == world:
            return t["traffic_bytes_per_launch"]
    except Exception:
        pass
    return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])
            if self.stop_flag.is_set():
                break
        self.proc.terminate()

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=2.0)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads() -> int:
    """Cores this process may use.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, which would
    quietly turn the CPU arm into a single-thread run; the oracle takes the thread count explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


LM_REF_CONFIGS = ("ladybug-49", "trafalgar-257")   # configs of BASELINE.json on which both arms time full LM runs
LM_REF_ITERS = 6


def workload_config(args, p):
    """The `config` object: identical in both arms."""
    return {"workload": args.workload, "ncams": p.ncams, "npnts": p.npnts, "nobs": p.nobs}


def reference_lm_legs(args, nt):
    """LM iters/s of the CPU restatement (src/benchmark.jl:31: Levenberg_Marquardt(FeasibilityResidual(BA), :LDL,
    :AMD, ...)): the oracle's src/lm.jl loop with the damped system eliminated in the order AMD gives it on a BA
    Jacobian (points first, dense Cholesky of the camera system), all host threads.  Whole runs of LM_REF_ITERS
    iterations on the small BASELINE.json configs; on the bench workload itself one iteration (evaluation + one
    exact solve, LAPACK for the dense factor) as the bounded sample."""
    import bundleadjustment.jl_b200.synth as synth
    from oracle import oracle as O
    out = {}
    for name in LM_REF_CONFIGS:
        q = synth.make_problem(name)
        prm = O.default_params(ite_max=LM_REF_ITERS - 1, nthreads=nt)
        t0 = time.perf_counter()
        r = O.lm_solve(q.cam_idx, q.pnt_idx, q.pt2d, q.ncams, q.npnts, q.x0, prm, solver="schur")
        dt = time.perf_counter() - t0
        out[name] = {"value": r.iter / dt, "unit": "LM iters/s", "iters": r.iter, "objective": r.objective,
                     "status": r.status, "seconds": dt}
    if args.lm_iters > 0:
        q = synth.make_problem(args.workload)
        if q.ncams <= 2048:
            t0 = time.perf_counter()
            O.lm_step_schur(q.cam_idx, q.pnt_idx, q.pt2d, q.ncams, q.npnts, q.x0, 30.0, nthreads=nt, dense="scipy")
            dt = time.perf_counter() - t0
            out[args.workload] = {"value": 1.0 / dt, "unit": "LM iters/s", "iters": 1, "seconds": dt,
                                  "sample": "one LM iteration: cons! + jac_coord! + J'r + one exact damped solve "
                                            "(Schur-ordered, LAPACK dpotrf on the %d x %d camera system)"
                                            % (9 * q.ncams, 9 * q.ncams)}
    return out


def run_reference(args):
    """CPU arm: the restatement of the reference (oracle/ba_oracle.c) on the host cores.  cons! uses all
    threads (Threads.@threads, src/BALNLPModels.jl:45); jac_coord! is capped at 3 like the reference
    (src/BALNLPModels.jl:167-168)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import bundleadjustment.jl_b200.synth as synth  # host-only generator
    from oracle import oracle as O
    p = synth.make_problem(args.workload)
    nt = host_threads()
    cx = np.empty(2 * p.nobs)
    vals = np.empty(24 * p.nobs)

    def step():
        O.lib().bao_cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, cx, p.nobs, p.npnts, nt)
        O.lib().bao_jac_coord(p.cam_idx, p.pnt_idx, p.x0, vals, p.nobs, p.npnts, min(nt, 3))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = p.nobs / dt / 1e6
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, p),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": nt, "kind": "port", "cpu_model": cpu_model(),
                            "sample": "full %s problem per step; cons! on %d threads, jac_coord! on %d "
                                      "(reference caps it at 3)" % (args.workload, nt, min(nt, 3))},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    if not args.no_lm_reference:
        out["lm_configs"] = reference_lm_legs(args, nt)
    print(json.dumps(out), flush=True)


def cpu_baseline(p, workload, budget_s=12.0):
    from oracle import oracle as O
    nt = host_threads()
    cx = np.empty(2 * p.nobs)
    vals = np.empty(24 * p.nobs)
    O.cons_jac(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts, nt, cx, vals)  # warm-up / page-in
    reps, t0 = 0, time.perf_counter()
    while True:
        O.cons_jac(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts, nt, cx, vals)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 50:
            break
    dt = (time.perf_counter() - t0) / reps
    n1 = min(p.nobs, 1_000_000)  # single-thread figure on a 1M-observation prefix of the same problem
    t1 = time.perf_counter()
    O.cons_jac(p.cam_idx[:n1], p.pnt_idx[:n1], p.pt2d[:2 * n1], p.x0, p.npnts, 1, cx[:2 * n1], vals[:24 * n1])
    dt1 = time.perf_counter() - t1
    return {"value": p.nobs / dt / 1e6, "unit": UNIT, "cores": nt, "kind": "port", "cpu_model": cpu_model(),
            "sample": "%d full passes of the %s problem (cons! + jac_coord!, all %d threads for both)"
                      % (reps, workload, nt),
            "single_thread_value": n1 / dt1 / 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="venice-1778")
    ap.add_argument("--lm-iters", type=int, default=8, help="LM iterations timed in the lm leg (0 = skip)")
    ap.add_argument("--pcg-max-iter", type=int, default=None, help="cap PCG iterations per solve (profiling runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lm-reference", action="store_true", help="skip the LM legs on the small configs")
    ap.add_argument("--solver", default="auto", choices=["auto", "pcg", "exact", "mixed"], help="damped solve of the LM leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import bundleadjustment.jl_b200 as ba
    L = ba._lib.lib()  # raises if the CUDA library is missing: no fallback

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    p = ba.synth.make_problem(args.workload)
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs, name=p.name,
                       device=local, rank=rank, nranks=world)
    h = m.handle
    nl = m.nobs_local
    # a side stream (the legacy default stream cannot be captured into the CUDA graph the PCG loop uses);
    # every torch.cuda.Event below is recorded on it, and the library launches on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ba._lib.check(L.ba_set_stream(h, C.c_void_p(stream.cuda_stream)), h)

    # ---- device-resident leg ---------------------------------------------------------------------
    x_d = torch.from_numpy(p.x0).cuda()
    step_bytes = 26 * 8 * nl
    ring = max(2, int(np.ceil(3 * 126e6 / max(step_bytes, 1))))  # outputs rotate through > 2x L2 of memory
    ring = min(ring, 16)
    cx_d = [torch.empty(2 * nl, dtype=torch.float64, device="cuda") for _ in range(ring)]
    vals_d = [torch.empty(24 * nl, dtype=torch.float64, device="cuda") for _ in range(ring)]

    def dev_step(i):
        j = i % ring
        rc = L.ba_residual_jac_dev(h, C.c_void_p(x_d.data_ptr()), C.c_void_p(cx_d[j].data_ptr()),
                                   C.c_void_p(vals_d[j].data_ptr()))
        if rc:
            ba._lib.check(rc, h)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(max(args.warmup, 3)):
        dev_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        dev_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev = float(t.item())

    # dominant kernel alone (k_eval<cx,vals>): the library brackets that launch with its own CUDA events
    # on the launching stream; read them step by step (same rotating buffers)
    ks = []
    ba._lib.check(L.ba_set_profiling(h, 1), h)
    for i in range(args.steps):
        dev_step(i)
        f = C.c_float()
        ba._lib.check(L.ba_last_eval_ms(h, C.byref(f)), h)
        ks.append(f.value)
    barrier()
    ba._lib.check(L.ba_set_profiling(h, 0), h)
    k_ms = float(np.mean(ks))

    # ---- jac_structure! once (src/lm.jl:53 calls it once per solve): write-only 384 B per observation -----------
    rows_d = torch.empty(24 * nl, dtype=torch.int64, device="cuda")
    cols_d = torch.empty(24 * nl, dtype=torch.int64, device="cuda")
    ba._lib.check(L.ba_jac_structure_dev(h, C.c_void_p(rows_d.data_ptr()), C.c_void_p(cols_d.data_ptr())), h)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(5):
        ba._lib.check(L.ba_jac_structure_dev(h, C.c_void_p(rows_d.data_ptr()), C.c_void_p(cols_d.data_ptr())), h)
    s1.record()
    barrier()
    js_ms = s0.elapsed_time(s1) / 5
    del rows_d, cols_d

    # ---- end-to-end leg through the host-pointer call (what Julia's ccall hits) -------------------------
    def pinned(nbytes):
        ptr = C.c_void_p()
        ba._lib.check(L.ba_alloc_pinned(nbytes, C.byref(ptr)))
        return ptr

    x_h, cx_h, vals_h = pinned(8 * p.nvar), pinned(16 * max(nl, 1)), pinned(192 * max(nl, 1))
    C.memmove(x_h, p.x0.ctypes.data, 8 * p.nvar)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ba._lib.check(L.ba_residual_jac(h, x_h, cx_h, vals_h), h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ba._lib.check(L.ba_residual_jac(h, x_h, cx_h, vals_h), h)   # returns with results on the host
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    h2d, d2h = 8 * p.nvar, 208 * nl
    for ptr in (x_h, cx_h, vals_h):
        L.ba_free_pinned(ptr)
    # the same call with ordinary (pageable) arrays -- what a plain `ccall` with an ordinary Vector{Float64} passes:
    # the library stages the copy through its own ring of pinned chunks, emptied by a few host threads
    cx_p, vals_p = np.empty(2 * max(nl, 1)), np.empty(24 * max(nl, 1))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    t0 = time.perf_counter()
    ba._lib.check(L.ba_residual_jac(h, vp(p.x0), vp(cx_p), vp(vals_p)), h)
    first_pageable_s = time.perf_counter() - t0
    ba._lib.check(L.ba_residual_jac(h, vp(p.x0), vp(cx_p), vp(vals_p)), h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ba._lib.check(L.ba_residual_jac(h, vp(p.x0), vp(cx_p), vp(vals_p)), h)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_pageable_s = float(t.item())
    del cx_p, vals_p

    # ---- LM leg: full Levenberg-Marquardt iterations per second on the same problem ----------------------
    lm, lm_configs, parity = None, None, None
    if args.lm_iters > 0:
        ba.init_comm(m)
        # untimed warm-up (the contract's W), identical at every N: a 3-iteration solve of a small synthetic problem
        # SHARDED like the timed one (own handle, own communicator), once per solver, so that every LM kernel, the
        # NCCL channels and the peer-memory exchange are loaded before the timed call.  Nothing of the timed problem
        # is precomputed: its schedules (lm_prepare) are built inside the timing.
        pw = ba.synth.make_problem((160, 10000, 50000))
        mw = ba.BALNLPModel(pw.cam_idx, pw.pnt_idx, pw.pt2d, pw.x0, pw.ncams, pw.npnts, pw.nobs, device=local,
                            rank=rank, nranks=world)
        ba._lib.check(L.ba_set_stream(mw.handle, C.c_void_p(stream.cuda_stream)), mw.handle)
        ba.init_comm(mw)
        steps_w = {}
        try:
            for sv in ("pcg", "exact", "mixed"):
                ba.Levenberg_Marquardt(mw, "LDL", "AMD", "None", False, ite_max=2, solver=sv,
                                       pcg_max_iter=args.pcg_max_iter)
                steps_w[sv] = ba.lm_step(mw, pw.x0, 30.0)[0]
        finally:
            mw.close()
        lm_warm = ("3-iteration solves (PCG, exact and mixed) of a (160, 10000, 50000) synthetic problem, sharded over the "
                   "same %d rank(s), on its own handle" % world)
        if world > 1 and rank == 0:
            # driver-visible multi-GPU parity: the sharded damped solve against the same solve on one GPU
            m1 = ba.BALNLPModel(pw.cam_idx, pw.pnt_idx, pw.pt2d, pw.x0, pw.ncams, pw.npnts, pw.nobs, device=local)
            parity = {"problem": "(160, 10000, 50000), lambda 30", "ranks": world}
            try:
                for sv in ("pcg", "exact", "mixed"):
                    m1.set_solver(sv)
                    d1 = ba.lm_step(m1, pw.x0, 30.0)[0]
                    parity[sv + "_rel_err_vs_1gpu"] = float(np.linalg.norm(steps_w[sv] - d1) / np.linalg.norm(d1))
                parity["ok"] = bool(max(parity[sv + "_rel_err_vs_1gpu"] for sv in ("pcg", "exact", "mixed")) <= 1e-10)
            finally:
                m1.close()
        barrier()
        t0 = time.perf_counter()
        st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=args.lm_iters - 1,
                                    pcg_max_iter=args.pcg_max_iter, solver=args.solver)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lm = {"metric": "LM iters/s", "value": st.iter / float(t.item()), "iters": st.iter,
              "solver": st.rows[0]["solver"] if st.rows else None,
              "solver_per_iteration": [r["solver"] for r in st.rows], "mixed_fallbacks": st.mixed_fallbacks,
              "pcg_iters": st.pcg_iters, "solver_iters_per_iteration": [r["pcg_iters"] for r in st.rows],
              "capped_solves": st.capped_solves, "worst_solve_rel": st.worst_solve_rel,
              "objective0": st.rows[0]["f"] if st.rows else None,
              "objective": st.objective, "status": st.status, "timings_ms": st.timings_ms,
              "e2e": "x0 host -> solution host through Levenberg_Marquardt(), schedules (lm_prepare) included",
              "warmup": lm_warm}
        if st.chol_count:
            fl = st.chol_n ** 3 / 3.0
            tf = fl * st.chol_count / (st.timings_ms["cholesky"] * 1e-3) / 1e12
            lm["cholesky"] = {"n": st.chol_n, "count": st.chol_count, "ms_each": st.timings_ms["cholesky"] / st.chol_count,
                              "TFLOPs": tf, "schur_assembly_ms_each": st.timings_ms["schur_assembly"] / st.chol_count,
                              "share_of_device_time": st.timings_ms["cholesky"] / st.timings_ms["device_total"]}
            pk = C.c_double(0.0)
            if lm["solver"] == "mixed" and st.mixed_fallbacks == 0:
                # dominant kernel of the mixed solver: the FP32 factorisation's trailing update on tcgen05 (kind::tf32,
                # three TF32 MMAs per FP32 product): executed TF32 flops against the TF32 tensor rate = half the
                # measured dense bf16 rate of MEASURED_PEAKS.json
                pkf = None
                try:
                    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                        pkf = float(json.load(f)["bf16_tflops"]) / 2.0
                except Exception:
                    pkf = 2250.0 / 2.0
                lm["roofline"] = {"bound": "tensor", "kernel": "ba::k_chol_syrk_tc (tcgen05.mma kind::tf32, 3 terms)",
                                  "achieved": 3.0 * tf, "peak": pkf * world, "unit": "TFLOP/s",
                                  "frac": 3.0 * tf / (pkf * world), "flops_per_factorisation": 3.0 * fl,
                                  "fp32_equivalent_TFLOPs": tf,
                                  "peak_source": "TF32 dense = MEASURED_PEAKS.json bf16_tflops / 2 (nominal 1125 if absent) "
                                                 "x %d rank(s)" % world}
            elif L.ba_measure_fp64_mma_peak(local, C.byref(pk)) == 0 and pk.value > 0:
                # dominant kernel of an LM iteration with the exact solver: the dense factorisation (k_chol_syrk, DMMA);
                # aggregate over the ranks when the factorisation is distributed
                lm["roofline"] = {"bound": "tensor", "kernel": "ba::k_chol_syrk (mma.sync.m8n8k4.f64)", "achieved": tf,
                                  "peak": pk.value * world, "unit": "TFLOP/s", "frac": tf / (pk.value * world),
                                  "flops_per_factorisation": fl,
                                  "peak_source": "ba_measure_fp64_mma_peak on this GPU x %d rank(s) (MEASURED_PEAKS.json "
                                                 "holds no FP64 figure)" % world}
        # a second, warm run on the same handle: the steady-state rate once the schedules exist
        barrier()
        t0 = time.perf_counter()
        st2 = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=args.lm_iters - 1,
                                     pcg_max_iter=args.pcg_max_iter, solver=args.solver)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lm["value_warm_handle"] = st2.iter / float(t.item())
        lm["objective_rerun_equal"] = bool(st2.objective == st.objective)
        if lm["solver"] in ("exact", "mixed"):
            # the other dense solver on the same (warm) handle: FP64 factorisation against FP32 factor + FP64 CG
            other = "mixed" if lm["solver"] == "exact" else "exact"
            for _rep in range(2):  # (the first call after a solver change rebuilds the LM state)
                barrier()
                t0 = time.perf_counter()
                st3 = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=args.lm_iters - 1,
                                             pcg_max_iter=args.pcg_max_iter, solver=other)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lm["other_dense_solver"] = {"solver": other, "value_warm_handle": st3.iter / float(t.item()),
                                        "objective": st3.objective, "timings_ms": st3.timings_ms,
                                        "solver_per_iteration": [r["solver"] for r in st3.rows],
                                        "solver_iters_per_iteration": [r["pcg_iters"] for r in st3.rows],
                                        "mixed_fallbacks": st3.mixed_fallbacks, "worst_solve_rel": st3.worst_solve_rel}
            m.set_solver(args.solver)
        if world == 1 and not args.no_lm_reference:
            # the configs the reference arm times whole LM runs on (bench.py --impl reference, lm_configs)
            lm_configs = {}
            for name in LM_REF_CONFIGS:
                q = ba.synth.make_problem(name)
                best = None
                for _rep in range(2):   # two fresh handles, the faster run is reported (device allocation times on a
                    mq = ba.BALNLPModel(q.cam_idx, q.pnt_idx, q.pt2d, q.x0, q.ncams, q.npnts, q.nobs, device=local)
                    try:                # shared host vary by tens of ms, which is the whole run at these sizes)
                        t0 = time.perf_counter()
                        sq = ba.Levenberg_Marquardt(mq, "LDL", "AMD", "None", False, ite_max=LM_REF_ITERS - 1)
                        dtq = time.perf_counter() - t0
                    finally:
                        mq.close()
                    if best is None or dtq < best[1]:
                        best = (sq, dtq)
                sq, dtq = best
                lm_configs[name] = {"value": sq.iter / dtq, "unit": "LM iters/s", "iters": sq.iter,
                                    "objective": sq.objective, "status": sq.status, "seconds": dtq,
                                    "solver": sq.rows[0]["solver"] if sq.rows else None,
                                    "sample": "whole run from x0, schedules and allocations included; best of 2 fresh handles"}
            lm_configs[args.workload] = {"value": lm["value"], "unit": "LM iters/s", "iters": lm["iters"]}

    if rank == 0:
        clocks = sampler.summary()  # sampled every 100 ms from warm-up to the end of the last leg
        peak, peak_src = measured_peak()
        bpo = algorithmic_bytes_per_obs(p)
        # per-launch algorithmic bytes of the dominant kernel on one rank (rank 0's shard)
        achieved = bpo * nl / (k_ms * 1e-3) / 1e9
        # secondary ceiling: FP64 FMA throughput measured on this GPU by the library's probe
        fp64 = C.c_double(0.0)
        fp64_peak = fp64.value if L.ba_measure_fp64_peak(local, C.byref(fp64)) == 0 else None
        flops_obs = 180.0  # algorithmic FP64 flops per observation with the per-camera precompute (section 8d)
        fp64_ach = flops_obs * nl / (k_ms * 1e-3) / 1e12
        out = {
            "metric": METRIC, "value": p.nobs / (ms_dev * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, p),
            "layout": {"sharding": "observations, contiguous point ranges, %d rank(s)" % world,
                       "l2": "per-step footprint %.0f MB, outputs rotate over %d buffers (> L2)" % (
                           step_bytes / 1e6, ring)},
            "e2e": {"value": p.nobs / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "pcie_GB/s": (h2d + d2h) / e2e_s / 1e9,
                    "note": "bound by the device-to-host copy of vals (192 B/obs) over PCIe, not by the kernel",
                    "api": "ba_residual_jac (host pointers), per rank, with the page-locked arrays the glue's allocating "
                           "cons()/jac_coord() hand to Levenberg_Marquardt (julia/BALNLPModels.jl: pinned_vector)",
                    "value_pageable_arrays": p.nobs / e2e_pageable_s / 1e6,
                    "pageable_note": "the same call with ordinary (pageable) numpy arrays: staged through the library's "
                                     "pinned ring by host threads (ba_hostio.cu)",
                    "first_pageable_call_s": first_pageable_s},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "kernel": "ba::k_eval<true,true,false>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(args.workload, world),
                         "traffic_unit": "bytes per launch (ncu dram read+write)", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bpo * nl,
                         "kernel_ms": k_ms, "bytes_per_obs": bpo, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "fp64": {"peak_TFLOPs_measured": fp64_peak, "flops_per_obs": flops_obs,
                                  "achieved_TFLOPs": fp64_ach, "frac": fp64_ach / fp64_peak if fp64_peak else None,
                                  "note": "secondary ceiling; ncu: FP64 pipe 21 % busy in k_eval"}},
            "clocks": clocks,
            "jac_structure": {"ms": js_ms, "GB/s": 392.0 * nl / (js_ms * 1e-3) / 1e9,
                              "bytes_per_obs": 392, "note": "rank 0 shard; 2 x 24 Int64 written + 8 B of indices read"},
        }
        if lm:
            out["lm"] = lm
        if lm_configs:
            out["lm_configs"] = lm_configs
        if parity:
            out["parity_multi_gpu"] = parity
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(p, args.workload)
        print(json.dumps(out), flush=True)
    m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
