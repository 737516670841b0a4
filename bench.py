#!/usr/bin/env python
"""Benchmark of the ELBO inner loop: negelcbo+grad evals/sec (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...    # the UNMODIFIED reference on the host CPU

A "step" is ONE evaluation of ``_neg_elcbo(theta, gp, vp, 0, Ns_K, compute_grad=True,
compute_var=False, theta_bnd)`` -- the ``minimize_adam`` objective
(pyvbmc/vbmc/variational_optimization.py:238-249) -- on the synthetic workload of SURVEY 8(d).

N = 1 : config C3 of BASELINE.json (D=20, N=400, K=50, S=8, N_s=400k  =>  8000 draws/component).
N > 1 : weak scaling, one C3 worth of entropy draws per GPU: N_s = 400k*N, S = max(8, 4*N)
        (N = 8 is exactly config C5: S=32, N_s=3.2M); draws and hyper-samples are sharded, one
        all-reduce of the raw (pre-Jacobian) vector per step.  ``value`` counts C3-equivalent
        evaluations (N per step) per second; ``evals_per_s_job`` is the plain number of
        (N-times larger) evaluations per second.

value : device-resident throughput (theta and GP already in HBM, CUDA events on the launching
        stream, L2 flushed between steps, max over ranks).  N = 1: the library's own launch
        sequence (``vbmc_negelcbo_enqueue``: the kernels the drop-in call runs, nothing else);
        N > 1: ``ShardedNegElcbo.enqueue`` (partials -> all-reduce -> finalize).
e2e   : the same evaluation through the reference-shaped public function with HOST NumPy
        buffers in and out (H2D of the parameters and D2H of (F, dF) inside the timed region).

Reference arm (``--impl reference``): the unmodified ``pyvbmc`` ``_neg_elcbo`` (from /root/reference,
or from the archive ``oracle/build_ref.py`` packs into the git-ignored ``oracle/_ref`` -- that is what
exists on the GPU box), FULL evaluations, nothing extrapolated.  Every step is one C3 evaluation at any
N (the host CPU does not grow with the GPU count; the reference's cost is linear in draws and
hyper-samples, so its C3-equivalent evaluations per second are the same number at every N).
``value`` is the faithful figure: one evaluation at a time, as the reference's optimiser calls it (NumPy's
BLAS threads where the path has any BLAS).  ``cpu_all_cores`` beside it: one single-threaded process per host
core, each running one full evaluation, started together -- what the host delivers on independent evaluations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "negelcbo+grad evals/sec at D=20,N=400,K=50,S=8,N_s=400k"
UNIT = "evals/s"
QUICK = os.environ.get("VBMC_BENCH_QUICK", "0") == "1"  # tests: skip the long CPU legs and the extras


def workload(n_gpus):
    from workloads import synthetic as syn

    S = 8 if n_gpus == 1 else max(8, 4 * n_gpus)
    pr = syn.make_problem("C3", S=S)
    pr.Ns_total = 400_000 * n_gpus
    pr.Ns_K = syn.ns_per_component(pr.Ns_total, pr.K)
    name = "C3 (D=20,N=400,K=50,S=8,N_s=400k)" if n_gpus == 1 else (
        f"C3 per GPU, weak: D=20,N=400,K=50,S={S},N_s={pr.Ns_total} sharded over {n_gpus} GPUs"
        + (" (= C5)" if n_gpus == 8 else "")
    )
    return pr, name


def flops_entmc(Ns_total, K, D):
    """Algorithmic flops of one entmc evaluation: N_s * [K (6D + 8) + 9D + 2]."""
    return float(Ns_total) * (K * (6 * D + 8) + 9 * D + 2)


def bytes_entmc(Ns_total, K, D, eps_input):
    """Algorithmic HBM bytes of one entmc launch: parameters in + one fp64 record per CTA out,
    plus the fp64 eps stream (N_s/2 * D * 8) when the draws are an input."""
    P = D * K + 2 * K + D
    b = 8.0 * (P + K)
    if eps_input:
        b += (Ns_total / 2) * D * 8.0
    return b


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the GPU is under the benchmark's load."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self._stop, self._t = device, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons,
            "samples": len(self.rows),
            "window": ("the untimed device-resident spin-up (same launch / flush pattern, GPU under load) plus the timed "
                       "region; the timed region alone lasts a few milliseconds, shorter than one nvidia-smi query"),
        }


# ------------------------------------------------------------------------------- CPU arms
def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return 1


class CpuArm:
    """FULL evaluations of the metric's call on the host CPU: the unmodified reference when it is available
    (``kind == "reference"``), else the oracle's NumPy restatement (``kind == "port"``)."""

    def __init__(self, pr, kind):
        self.pr, self.kind = pr, kind
        if kind == "reference":
            from oracle import ref_loader

            self.ref = ref_loader.load()
            self.gp = ref_loader.make_ref_gp(pr.X, pr.y, pr.posts, pr.mean_kind)
            self.vp = ref_loader.make_ref_vp(pr.D, pr.K, pr.mu, pr.sigma, pr.lambd, pr.w, pr.eta, pr.optimize)
            self.where = self.ref._neg_elcbo.__code__.co_filename
        else:
            from oracle import elbo_oracle as eo

            self.eo = eo
            self.gp = pr.gp
            self.vp = eo.OracleVP.create(pr.D, pr.K, pr.mu, pr.sigma, pr.lambd, pr.w, pr.eta, pr.optimize)
            self.where = eo.__file__

    @staticmethod
    def best_kind():
        if os.environ.get("VBMC_BENCH_CPU", "") == "port":
            return "port"
        try:
            from oracle import ref_loader

            return "reference" if ref_loader.available() else "port"
        except Exception:
            return "port"

    def evaluate(self, Ns_K=None):
        """One full ``_neg_elcbo(theta, gp, vp, 0, Ns_K, True, False, theta_bnd)``; the draws come from the global
        NumPy stream inside the call, exactly as in the reference (entmc_vbmc.py:64-68)."""
        pr = self.pr
        Ns_K = pr.Ns_K if Ns_K is None else Ns_K
        if self.kind == "reference":
            return self.ref._neg_elcbo(pr.theta.copy(), self.gp, self.vp, 0.0, Ns_K, True, False, pr.theta_bnd)
        return self.eo.neg_elcbo(pr.theta.copy(), self.gp, self.vp, 0.0, Ns_K, True, False, pr.theta_bnd)

    def time(self, steps, warmup):
        """-> (seconds per evaluation [mean of the timed steps], per-step list, cores kept busy, wall of timed part)."""
        np.random.seed(0)
        for _ in range(warmup):
            self.evaluate()
        ts = []
        cpu0, wall0 = time.process_time(), time.perf_counter()
        for _ in range(steps):
            t0 = time.perf_counter()
            F, dF, *_ = self.evaluate()
            ts.append(time.perf_counter() - t0)
        wall = time.perf_counter() - wall0
        cores = max(1.0, round((time.process_time() - cpu0) / max(wall, 1e-9), 1))
        assert np.isfinite(F) and np.all(np.isfinite(dF))
        return float(np.mean(ts)), ts, cores, wall

    def describe(self, steps, warmup, wall):
        what = ("UNMODIFIED reference pyvbmc._neg_elcbo (" + self.where + ")" if self.kind == "reference"
                else "oracle port (fp64 NumPy restatement of the reference, oracle/elbo_oracle.py)")
        pr = self.pr
        return (f"{what}; {steps} FULL evaluations of C3 (all {pr.Ns_K} draws x {pr.K} components, all {pr.S} "
                f"hyper-samples, gradient, soft bounds) after {warmup} full warm-up evaluation(s); nothing sampled or "
                f"extrapolated; {wall:.1f} s wall")


def _cpu_worker(start_at):
    """One process of the all-cores figure: set up, warm up on a small draw count, wait for the common start time, run ONE
    full evaluation, print its duration."""
    pr, _ = workload(1)
    arm = CpuArm(pr, CpuArm.best_kind())
    np.random.seed(os.getpid() % 65536)
    arm.evaluate(Ns_K=40)
    late = time.time() - start_at
    if late < 0:
        time.sleep(-late)
    t0 = time.perf_counter()
    F, dF, *_ = arm.evaluate()
    dt = time.perf_counter() - t0
    print(json.dumps({"worker_s": dt, "late_s": max(late, 0.0), "ok": bool(np.isfinite(F))}))
    return 0


def cpu_all_cores(n_proc, setup_s=30.0):
    """The same evaluation on every host core at once: `n_proc` independent processes (one BLAS thread each), each running
    ONE full evaluation of the unmodified reference, started together; throughput = n_proc / slowest process.  The
    reference itself offers no parallelism over one evaluation (NumPy elementwise kernels in a Python loop over the
    components), so independent evaluations -- e.g. the candidates of the sieve -- are what extra cores can take."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    start_at = time.time() + setup_s
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--cpu-worker", repr(start_at)],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env, text=True) for _ in range(n_proc)]
    outs = []
    for p_ in procs:
        try:
            o, _ = p_.communicate(timeout=600)
            outs.append(json.loads(o.strip().splitlines()[-1]))
        except Exception:
            p_.kill()
    if len(outs) < n_proc or not all(o["ok"] for o in outs) or any(o["late_s"] > 5.0 for o in outs):
        return {"unavailable": f"{len(outs)} of {n_proc} workers finished cleanly / on time"}
    slow = max(o["worker_s"] for o in outs)
    return {"value": n_proc / slow, "unit": UNIT, "processes": n_proc, "slowest_s": slow,
            "fastest_s": min(o["worker_s"] for o in outs),
            "what": ("independent FULL evaluations of the unmodified reference, one single-threaded process per host core, "
                     "started together; value = processes / slowest process")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.cpu_worker is not None:
        return _cpu_worker(float(args.cpu_worker))
    pr, _ = workload(1)  # one C3 evaluation per step at every N (see the module docstring)
    kind = CpuArm.best_kind()
    arm = CpuArm(pr, kind)
    t_eval, ts, cores, wall = arm.time(args.steps, args.warmup)
    all_cores = cpu_all_cores(os.cpu_count() or 1) if os.environ.get("VBMC_BENCH_ALL_CORES", "1") != "0" else None
    value = 1.0 / t_eval  # C3 evaluations per second == C3-equivalent evaluations per second of the CUDA arm's unit
    wname = "C3 (D=20,N=400,K=50,S=8,N_s=400k)"
    if args.gpus > 1:
        wname += (f"; the CUDA arm at {args.gpus} GPUs evaluates {args.gpus} C3-equivalents per step -- the host CPU does "
                  "not grow with the GPU count and the reference's cost is linear in draws and hyper-samples, so one "
                  "step here is one full C3 evaluation (1/N of the N-GPU step) and `value` is in the same unit")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_eval, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wname, "timing": "host wall clock (time.perf_counter) around every evaluation",
                   "ms_per_step_min": 1e3 * float(np.min(ts)), "ms_per_step_max": 1e3 * float(np.max(ts))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": arm.describe(args.steps, args.warmup, wall), "host_cpus": os.cpu_count(),
                         "blas_threads": blas_threads(),
                         "threads_note": ("NumPy's elementwise kernels (97 % of this path) are single-threaded; only the "
                                          "np.dot / solve_triangular calls fan out over the BLAS threads; `cores` = process "
                                          "CPU time / wall time")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if all_cores is not None:
        line["cpu_all_cores"] = all_cores
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------- CUDA arm
def _make_vp(pv, pr):
    vp = pv.VariationalPosterior(pr.D, pr.K)
    vp.mu, vp.sigma, vp.lambd = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy()
    vp.w, vp.eta = pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()
    return vp


def _spin(step_fn, world, fixed_blocks, block, budget_s=3.0):
    """Untimed spin-up: the GPU sits in a low-power state after the imports and needs a while under load before
    its clocks settle.  One GPU: until two consecutive blocks agree to 3 % (at most ``budget_s``).  N ranks: every
    rank must run the SAME number of evaluations (each holds an all-reduce), so a fixed count.  Returns the number of
    untimed evaluations run."""
    t0, prev, n = time.perf_counter(), None, 0
    for blk_i in range(100):
        b = step_fn(block)
        n += block
        if world == 1:
            if (prev is not None and abs(b - prev) <= 0.03 * prev and time.perf_counter() - t0 > 0.5) or \
                    time.perf_counter() - t0 > budget_s:
                break
        elif blk_i + 1 >= fixed_blocks:
            break
        prev = b
    return n


def _extras_n1(pv, pr, torch, flush):
    """Informational lines for the next rows of SURVEY 8(f) and the smaller configs (NOT part of value / e2e)."""
    from workloads import synthetic as syn

    ex = {}

    def small(name, prs, Ns_K, reps=200):
        vp = _make_vp(pv, prs)
        call = lambda: pv._neg_elcbo(prs.theta, prs.gp, vp, 0.0, Ns_K, True, False, prs.theta_bnd)  # noqa: E731
        for _ in range(300):
            call()
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        e2e_us = 1e6 * (time.perf_counter() - t0) / reps
        ctx = pv.context_for_gp(prs.gp)
        st = torch.cuda.ExternalStream(ctx.stream)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(20):
            ctx.enqueue()
        ctx.synchronize()
        a.record(st)
        for _ in range(reps):
            ctx.enqueue()
        b.record(st)
        b.synchronize()
        ex[name] = {"e2e_us_per_eval": e2e_us, "device_us_per_eval": 1e3 * a.elapsed_time(b) / reps,
                    "draws_per_component": Ns_K, "D": prs.D, "K": prs.K, "N": prs.N, "S": prs.S,
                    "what": "same call as the headline (value + gradient, soft bounds), warm L2, back to back"}

    small("C3_back_to_back", pr, pr.Ns_K)  # the headline workload without the L2 flush between steps
    small("C2", syn.make_problem("C2"), syn.make_problem("C2").Ns_K)
    small("C4", syn.make_problem("C4"), syn.make_problem("C4").Ns_K)
    small("C3_at_28_draws_per_component", pr, 28)

    # device-resident Adam (N2)
    vpa = _make_vp(pv, pr)
    th0 = np.asarray(vpa.get_parameters(), dtype=float)
    kw = dict(seed=1, max_iter=200, use_early_stopping=False, master_max=0.01)
    pv.minimize_adam_elcbo(pr.gp, vpa, th0.copy(), pr.Ns_K, pr.theta_bnd, **kw)  # warm-up (graph capture, clocks)
    t0 = time.perf_counter()
    _, _, _, yt, n_it = pv.minimize_adam_elcbo(pr.gp, vpa, th0.copy(), pr.Ns_K, pr.theta_bnd, **kw)
    dt = time.perf_counter() - t0
    ex["device_adam"] = {
        "iterations_per_s": n_it / dt, "us_per_iteration": 1e6 * dt / n_it, "iterations": int(n_it),
        "what": ("pyvbmc_b200.minimize_adam_elcbo: minimize_adam.py:61-145 with theta, moments and iterates resident in "
                 "HBM, one CUDA graph per iteration (evaluation + Adam update), host sync every 20 iterations; same "
                 "workload as `value` (one negelcbo+grad evaluation per iteration)"),
    }
    # batched sieve (N1)
    Bs = 500
    rng = np.random.default_rng(0)
    cands = []
    for _ in range(Bs):
        v = _make_vp(pv, pr)
        v.mu = pr.mu + 0.3 * rng.normal(size=pr.mu.shape)
        cands.append(v)
    pv.neg_elcbo_batch(cands, pr.gp, pr.theta_bnd)  # warm-up (clocks, kernel attributes)
    t0 = time.perf_counter()
    pv.neg_elcbo_batch(cands, pr.gp, pr.theta_bnd)
    dt = time.perf_counter() - t0
    ex["sieve_batch"] = {
        "candidates_per_s": Bs / dt, "us_per_candidate": 1e6 * dt / Bs, "candidates": Bs,
        "what": ("pyvbmc_b200.neg_elcbo_batch: the value-only candidate loop of variational_optimization.py:775-787 "
                 "(entlb + log joint + bounds) as one launch, host packing included"),
    }
    # variance path (N3): _eval_full_elcbo's call, value + variance + per-component terms
    try:
        vpv = _make_vp(pv, pr)
        callv = lambda: pv._neg_elcbo(pr.theta, pr.gp, vpv, 0.0, 0, False, True, None, 0.0, True)  # noqa: E731
        for _ in range(3):
            callv()
        t0 = time.perf_counter()
        for _ in range(10):
            callv()
        us = 1e6 * (time.perf_counter() - t0) / 10
        N, S, K = pr.N, pr.S, pr.K
        b_alg, f_alg = S * N * N * 8.0, S * (2.0 * N * N * K + 2.0 * K * K * N)
        ex["variance_path"] = {
            "us_per_call": us, "algorithmic_bytes": b_alg, "algorithmic_flops": f_alg,
            "hbm_GBps_achieved": b_alg / (us * 1e-6) / 1e9, "fp64_TFLOPs_achieved": f_alg / (us * 1e-6) / 1e12,
            "what": ("_neg_elcbo(theta, gp, vp, 0, 0, False, True, None, 0, separate_K=True): _gp_log_joint variance path "
                     "(variational_optimization.py:1472-1518) end to end through the drop-in call, deterministic entropy"),
        }
    except Exception as exc:
        ex["variance_path"] = {"error": repr(exc)}
    # acquisition-function ingredients (N4): GP prediction at the reference's search-cache size, mixture density
    try:
        rng = np.random.default_rng(0)
        Nx = 8192
        Xs = pr.X[rng.integers(0, pr.N, size=Nx)] + 0.5 * rng.normal(size=(Nx, pr.D))
        pv.gp_predict(pr.gp, Xs, separate_samples=True)  # warm-up: packs L, builds the triangular inverses
        t0 = time.perf_counter()
        for _ in range(5):
            pv.gp_predict(pr.gp, Xs, separate_samples=True)
        e2e_ms = 1e3 * (time.perf_counter() - t0) / 5
        ctxL = pv.context_for_gp(pr.gp, need_L=True)
        dev_ms = ctxL.gp_predict_device_ms(Nx, 10)
        N, S = pr.N, pr.S
        f_alg = float(S) * Nx * (N * N + 2.0 * N)  # triangular product W = K* M (N^2 Nx flops) + the two dot products
        vpp = _make_vp(pv, pr)
        pv.vp_pdf(vpp, Xs, orig_flag=False, log_flag=True)
        t0 = time.perf_counter()
        for _ in range(5):
            pv.vp_pdf(vpp, Xs, orig_flag=False, log_flag=True)
        pdf_ms = 1e3 * (time.perf_counter() - t0) / 5
        from oracle import acq_oracle as ao

        sub = 512
        t0 = time.perf_counter()
        ao.gp_predict(pr.X, pr.posts, Xs[:sub], pr.mean_kind, separate_samples=True)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        ex["gp_predict"] = {
            "points": Nx, "e2e_ms": e2e_ms, "device_ms": dev_ms, "points_per_s_e2e": Nx / (e2e_ms * 1e-3),
            "algorithmic_flops": f_alg, "fp64_TFLOPs_achieved": f_alg / (dev_ms * 1e-3) / 1e12,
            "vp_pdf_e2e_ms": pdf_ms,
            "cpu_oracle_ms_per_512_points": cpu_ms,
            "what": ("pyvbmc_b200.gp_predict(gp, Xs, separate_samples=True) at the reference's search-cache size "
                     "(abstract_acq_fcn.py:79); device_ms = gppred_kernel alone (DMMA fp64), e2e includes the H2D of Xs "
                     "and the D2H of f_mu / f_s2; the CPU figure is the oracle's NumPy restatement of gpyreg's predict on "
                     "512 of the points (its cost is linear in the number of points)"),
        }
    except Exception as exc:
        ex["gp_predict"] = {"error": repr(exc)}
    # the drop-in claim, timed: the UNMODIFIED reference optimize_vp (variational_optimization.py:90-391) on the C2 GP,
    # unpatched (CPU) and on top of pyvbmc_b200.install(device_adam=True, batched_sieve=True)
    try:
        from oracle import ref_loader

        if ref_loader.available():
            ref = ref_loader.load()
            import pyvbmc.vbmc as vpk
            from pyvbmc.vbmc import variational_optimization as vo
            from pyvbmc.vbmc.options import Options

            p2 = syn.make_problem("C2")
            base = os.path.join(os.path.dirname(vpk.__file__), "option_configs")
            opts = Options(os.path.join(base, "basic_vbmc_options.ini"), evaluation_parameters={"D": p2.D},
                           user_options={"max_iter_stochastic": 200})
            opts.load_options_file(os.path.join(base, "advanced_vbmc_options.ini"), evaluation_parameters={"D": p2.D})
            gp2 = ref_loader.make_ref_gp(p2.X, p2.y.reshape(-1, 1), p2.posts, p2.mean_kind)

            def once():
                np.random.seed(0)
                v = ref_loader.make_ref_vp(p2.D, p2.K, p2.mu, p2.sigma, p2.lambd, p2.w, p2.eta)
                t0 = time.perf_counter()
                v2, _, _ = vo.optimize_vp(opts, {"warmup": False, "entropy_switch": False}, v, gp2, 100, 1, p2.K)
                return time.perf_counter() - t0, float(v2.stats["elbo"])

            t_ref, elbo_ref = once()
            pv.install(device_adam=True, batched_sieve=True)
            try:
                once()  # warm-up (graph capture)
                t_dev, elbo_dev = once()
            finally:
                pv.uninstall()
            ex["dropin_optimize_vp"] = {
                "unpatched_s": t_ref, "patched_s": t_dev, "speedup": t_ref / t_dev, "elbo_unpatched": elbo_ref,
                "elbo_patched": elbo_dev,
                "what": ("unmodified pyvbmc optimize_vp on the C2 GP (D=10, N=200, K=20, S=4; 100 sieve candidates, 1 slow "
                         "start, <= 200 Adam iterations at the reference's default draw count, full-ELCBO evaluation, "
                         "pruning trials): wall time on the reference's NumPy functions vs on pyvbmc_b200.install("
                         "device_adam=True, batched_sieve=True)"),
            }
    except Exception as exc:
        ex["dropin_optimize_vp"] = {"error": repr(exc)}
    return ex


def run_b200(args):
    import torch
    import torch.distributed as dist

    import pyvbmc_b200 as pv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pv.config.device = local

    pr, wname = workload(world)
    D, K = pr.D, pr.K
    vp = _make_vp(pv, pr)
    W = max(args.warmup, 3)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = None
    p2p = False
    if world == 1:
        # the drop-in function itself, and the library's own launch sequence for the device-resident steps
        ctx = pv.context_for_gp(pr.gp)

        def e2e_step():
            return pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, pr.Ns_K, True, False, pr.theta_bnd)

        enqueue = ctx.enqueue
    else:
        from pyvbmc_b200.distributed import ShardedNegElcbo

        ev = ShardedNegElcbo(pr.gp, device=local, seed=1234)
        # raw-vector all-reduce: over NVLink peer memory inside the tail kernel when every rank can map every peer
        # (negotiated collectively, checked after the first evaluation), else NCCL.  VBMC_BENCH_P2P=0 forces NCCL.
        if os.environ.get("VBMC_BENCH_P2P", "1") == "1":
            p2p = ev.enable_p2p(D, K)
        ctx = ev.ctx

        def e2e_step():
            return ev(pr.theta, vp, pr.Ns_K, pr.theta_bnd)

        def enqueue():
            ev.enqueue(D, K)

    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    F0 = e2e_step()[0]  # stages theta / bounds for the device-resident loop as well
    if p2p:
        p2p = ev.p2p_self_check(F0)  # a timed-out peer exchange on any rank => NCCL all-reduce on all ranks

    # ---- e2e: public function, host buffers ----------------------------------------------------------------
    for _ in range(W):
        F, dF, G, H, _ = e2e_step()
    assert np.isfinite(F) and np.all(np.isfinite(dF))

    def e2e_block(n):
        b0 = time.perf_counter()
        for _ in range(n):
            e2e_step()
        return time.perf_counter() - b0

    spin_e2e = _spin(e2e_block, world, fixed_blocks=10, block=50)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    lay_total = K * D + 5 * K + 2 * D
    P = D * K + 2 * K + D
    h2d = 8 * (lay_total + 2)
    d2h = 8 * (8 + P)

    # ---- device-resident steps: CUDA events per step on the launching stream --------------------------------
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for _ in range(W):
        enqueue()
    barrier()
    wa, wb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def dev_block(n):
        acc = 0.0
        for i in range(n):
            with torch.cuda.stream(stream):
                flush.fill_(float(i))
            wa.record(stream)
            enqueue()
            wb.record(stream)
            wb.synchronize()
            acc += wa.elapsed_time(wb)
        return acc

    with ClockSampler(local) as clk:
        spin_dev = _spin(dev_block, world, fixed_blocks=15, block=20)
        barrier()
        launches1 = ctx.launch_count
        for i in range(args.steps):
            # evict L2 on the evaluation's own stream: stream order keeps the fill out of the [start, stop] interval
            # without a host synchronisation per step (with N ranks a per-step host barrier only measures launch skew;
            # the ranks stay in lock-step through the all-reduce of every step)
            with torch.cuda.stream(stream):
                flush.fill_(float(i))
            starts[i].record(stream)
            enqueue()
            stops[i].record(stream)
        barrier()
    launches_timed = ctx.launch_count - launches1
    ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    t_step = float(np.mean(ms)) * 1e-3
    if world > 1:
        t = torch.tensor([t_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step = float(t.item())

    # ---- roofline of the dominant kernel (entmc), per-launch CUDA events on its stream ----------------------
    ctx.set_kernel_timing(True)
    for _ in range(3):
        enqueue()
    ctx.synchronize()
    ctx.entmc_kernel_ms()
    for i in range(min(args.steps, 20)):
        with torch.cuda.stream(stream):
            flush.fill_(float(i))
        enqueue()
    ctx.synchronize()
    k_main_ms = ctx.entmc_main_kernel_ms()
    k_ms, k_n = ctx.entmc_kernel_ms()
    if k_main_ms <= 0.0:  # another kernel variant than the tensor-core one: a single launch
        k_main_ms = k_ms
    ctx.set_kernel_timing(False)
    barrier()

    # ---- N > 1: the sharded result against ONE GPU evaluating the whole job on the same Philox key ----------
    parity = None
    if world > 1:
        from pyvbmc_b200.distributed import ShardedNegElcbo

        single = ShardedNegElcbo(pr.gp, device=local, seed=1234, single=True)
        ev.step = single.step = 100
        Fs, dFs, Gs, Hs, _ = ev(pr.theta, vp, pr.Ns_K, pr.theta_bnd)
        F1, dF1, G1, H1, _ = single(pr.theta, _make_vp(pv, pr), pr.Ns_K, pr.theta_bnd)
        single.close()
        t = torch.tensor([Fs], dtype=torch.float64, device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        parity = {"rel_F": abs(Fs - F1) / abs(F1), "rel_G": abs(Gs - G1) / abs(G1), "rel_H": abs(Hs - H1) / abs(H1),
                  "rel_dF": float(np.abs(dFs - dF1).max() / np.abs(dF1).max()),
                  "replicated_bitwise": bool(float(lo) == float(hi)),
                  "what": "sharded evaluation vs ONE GPU evaluating the whole N-GPU job, same Philox key"}

    extras = None
    if world == 1 and not QUICK:
        try:
            extras = _extras_n1(pv, pr, torch, flush)
        except Exception as exc:  # extras must never take the headline down
            extras = {"error": repr(exc)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        fp32_peak = max(ctx.fma_peak(0), ctx.fma_peak(2))  # scalar FFMA vs packed FFMA2, whichever is higher
        fp64_peak = ctx.fma_peak(1)
        Ns_rank = pr.Ns_total / world
        b_alg = bytes_entmc(Ns_rank, K, D, eps_input=False)
        f_alg = flops_entmc(Ns_rank, K, D)
        k_s = k_main_ms * 1e-3
        achieved_gbs = b_alg / k_s / 1e9
        variant = ctx.entmc_variant_used()
        kname = {5: "entmc_kernel_tc<20,ANYGRAD> (tcgen05 / TMEM)",
                 4: "entmc_kernel_w<20,WGRAD,ANYGRAD,PHILOX>", 0: "entmc_kernel_fast<20,...>"}.get(variant, f"entmc variant {variant}")
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "entmc_traffic.json")))
            traffic = tj.get(f"variant{variant}", {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {
            "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
            "traffic": traffic, "kernel": kname, "entmc_variant": variant, "kernel_ms": k_main_ms,
            "kernel_ms_with_table_kernel": k_ms,
            "kernel_launches_timed": k_n, "algorithmic_bytes_per_launch": b_alg, "peak_source": hbm_src,
            "timing": ("CUDA events on the launching stream around the dominant kernel alone (`kernel_ms`) and around the "
                       "table kernel + dominant kernel (`kernel_ms_with_table_kernel`); one host synchronisation per timed "
                       "launch, L2 flushed before each.  The noise generator of the NEXT evaluation runs on a side stream "
                       "beside the tail kernel and is outside both brackets"),
            "traffic_note": ("dram__bytes of one entmc_kernel_tc launch under `ncu --set full` (profiles/r2x_ncu_summary.md): "
                             "ncu flushes L2 before the launch, so the 40 MB of noise-tile images the generator left in L2 "
                             "are re-read from DRAM there; in the timed loop they are L2 hits"),
            "note": ("entmc is bound by instruction issue (FP32 FMA / MUFU), not HBM: arithmetic intensity >> ridge; with "
                     "device Philox draws its only ALGORITHMIC HBM traffic is the parameter block and one record per CTA, "
                     "so the mandated HBM fraction is tiny by construction; the binding roofline is `compute`."),
            "compute": {
                "bound": "fp32_fma", "achieved": f_alg / k_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": f_alg / k_s / 1e12 / fp32_peak if fp32_peak else None,
                "algorithmic_flops_per_launch": f_alg,
                "peak_source": "max(FFMA, FFMA2) issue peak measured by vbmc_fma_peak in this run", "fp64_fma_peak": fp64_peak,
            },
        }
        cpu = None
        if world == 1:
            kind = "port" if QUICK else CpuArm.best_kind()
            arm = CpuArm(pr, kind)
            n_cpu = 1 if QUICK else 3
            t_cpu, ts_cpu, cores, wall = arm.time(n_cpu, 0 if QUICK else 1)
            cpu = {"value": 1.0 / t_cpu, "unit": UNIT, "cores": cores, "kind": kind, "blas_threads": blas_threads(),
                   "sample": arm.describe(n_cpu, 0 if QUICK else 1, wall), "host_cpus": os.cpu_count()}
            if kind == "reference":  # the oracle port beside it (it is what the parity tests run on the box)
                t_port, _, cores_p, wall_p = CpuArm(pr, "port").time(2, 1)
                cpu["port_beside_it"] = {"value": 1.0 / t_port, "unit": UNIT, "cores": cores_p,
                                         "sample": f"oracle port, 2 full C3 evaluations after 1 warm-up, {wall_p:.1f} s wall"}
        line = {
            "metric": METRIC, "value": world / t_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (entropy kernel: fp32 compute, fp64 accumulation) + f64 (log-joint, finalize)",
            "data": "synthetic",
            "config": {"workload": wname, "rng": "device Philox4x32-10 + Box-Muller",
                       "l2": "flushed between steps (256 MiB fill on the same stream, outside the timed interval)",
                       "warmup_steps_run": W,
                       "spinup_steps_untimed": {"e2e": spin_e2e, "device_resident": spin_dev,
                                                "why": "GPU clock settling after start-up: blocks repeated until two agree to 3 % (<= 3 s); fixed count with N ranks"},
                       "timing": "CUDA events per step on the launching stream, max over ranks",
                       "value_path": ("vbmc_negelcbo_enqueue (the library's own kernels of one drop-in call, parameters resident)"
                                      if world == 1 else "ShardedNegElcbo.enqueue (partials -> all-reduce -> finalize)"),
                       "draws_per_component": pr.Ns_K, "S": pr.S, "parallelism": f"draws+hyper-samples sharded x{world}",
                       "all_reduce": ("none" if world == 1 else ("peer memory (NVLink P2P stores + flags) inside the tail kernel"
                                                                 if p2p else "NCCL"))},
            "evals_per_s_job": 1.0 / t_step,
            "e2e": {"value": world / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e,
                    "path": ("pyvbmc_b200._neg_elcbo(theta, gp, vp, 0, Ns_K, True, False, theta_bnd)" if world == 1 else
                             "pyvbmc_b200.distributed.ShardedNegElcbo.__call__") + " (NumPy theta in -> NumPy (F, dF) out)"},
            "gpu_launches": int(launches_timed),
            "gpu_launches_per_step": launches_timed / args.steps,
            "clocks": clk.summary(),
            "roofline": roofline,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if parity:
            line["parity"] = parity
        if extras:
            line["extras"] = extras
        print(json.dumps(line))
    torch.cuda.synchronize()
    if ev is not None:
        ev.close()
    pv.clear_caches()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-worker", default=None, help=argparse.SUPPRESS)  # internal: one process of cpu_all_cores()
    args = ap.parse_args()
    # defaults that finish within minutes: a CUDA step lasts ~0.1 ms, a full reference evaluation several seconds
    if args.steps is None:
        args.steps = 50 if args.impl == "b200" else 10
    if args.warmup is None:
        args.warmup = 5 if args.impl == "b200" else 1
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
