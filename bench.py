#!/usr/bin/env python
"""Benchmark of the per-trial acquisition sweep (BASELINE.json metric: EI grid-points/s per trial, all
exploration sets) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sets-per-gpu S] [--strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload: BASELINE.json configs[4], the synthetic scaled sweep (1e6-point intervention grid per exploration
set, 1e4 observational samples, 32 interventional rows, d = 3 + 3 conditioning columns),
weak-scaled: every GPU sweeps `--sets-per-gpu` (default 2) exploration sets, so 8 GPUs run exactly the
16-set configuration.  `--strong` sweeps the full 16 sets on however many GPUs there are.
A step = one full post-observation trial: exp tables, prior precompute, prior on x_int and on the grid,
posterior fit, EI / cost, argmax, and (N > 1) the NCCL all-gather + combine of the per-set bests.

One JSON line on stdout (rank 0).  `--impl reference` times the CPU restatement of the reference's
arithmetic (oracle/, the reference itself cannot be installed offline -- DESIGN.md) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "EI grid-points/sec per trial (all exploration sets)"
UNIT = "grid-points/s"
N_OBS, P_GRID, D_INT, C_COND, N_INT = 10_000, 100, 3, 3, 32
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r01_prior_eval_traffic.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sets-per-gpu", type=int, default=2)
    ap.add_argument("--strong", action="store_true", help="fixed total work: the full 16-set sweep split over the GPUs")
    ap.add_argument("--n-obs", type=int, default=N_OBS)
    ap.add_argument("--p", type=int, default=P_GRID)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    return ap.parse_args()


def workload_config(args, world):
    S = 16 if args.strong else args.sets_per_gpu * world
    return {
        "workload": "synthetic scaled sweep (BASELINE.json configs[4]): %d exploration sets x %d^3 grid points x %d "
                    "observational samples, d=3 intervened + 3 conditioning columns, %d interventional rows%s"
                    % (S, args.p, args.n_obs, N_INT, "" if args.strong else
                       " -- weak-scaled, %d sets per GPU (16 sets at 8 GPUs is the full configuration)" % args.sets_per_gpu),
        "exploration_sets": S, "grid_points_per_set": args.p ** D_INT, "n_obs": args.n_obs, "n_int": N_INT,
        "step": "full post-observation trial: tables + prior precompute + prior(x_int) + posterior fit + prior(grid) + EI/cost + argmax"
                + (" + NCCL all-gather/combine" if world > 1 else ""),
        "l2": "inputs larger than L2: every set streams its 0.8 GB prior matrix M (L2 is 126 MB); no explicit flush",
        "partition": "contiguous FLOP-weighted chunks of (set x grid tile), one per GPU (cbo_with_oop_b200/partition.py)",
    }, S


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/cbo_oracle.py) on the host cores.  Only this function touches oracle/.
# ---------------------------------------------------------------------------------------------------------
def cpu_port(problem, sample_pts, seed=0, direct_pts=0):
    """Time the best-effort vectorised CPU port (factorised prior + batched posterior/EI in NumPy/BLAS) on a
    bounded sample of one exploration set and extrapolate linearly to the set's full grid."""
    from oracle import cbo_oracle as O
    try:
        from threadpoolctl import threadpool_info
        blas_threads = max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:
        blas_threads = os.cpu_count() or 1
    X = np.hstack([problem.x_obs_int, problem.x_obs_cond])
    d = problem.d
    N = X.shape[0]
    ls = np.concatenate([problem.ls_int, problem.ls_cond])
    gp = dict(X=X, variance=problem.s2, lengthscale=ls, noise=problem.noise, alpha=problem.alpha_obs, Kyinv=problem.kyinv,
              form="diff")
    cols = list(range(d))
    t0 = time.perf_counter()
    factors = O.prior_factors(gp, X, cols)                      # one-off per set (K1a's work)
    mI, vI = O.do_prior_factorised(gp, factors, cols, problem.x_int)
    post = O.posterior_fit(problem.x_int, problem.y_int, mI, vI, form="diff")
    t_once = time.perf_counter() - t0
    rng = np.random.default_rng(seed)
    G = problem.g_total
    flat = np.sort(rng.choice(G, size=min(sample_pts, G), replace=False))
    ii = np.unravel_index(flat, [len(t) for t in problem.grid])
    Xg = np.stack([problem.grid[k][ii[k]] for k in range(d)], axis=1)
    best = float(problem.y_int.min())
    t0 = time.perf_counter()
    mg, vg = O.do_prior_factorised(gp, factors, cols, Xg, chunk=2048)
    mu, var = O.posterior_predict(post, Xg, mg, vg)
    acq = O.expected_improvement(mu, var, best, "min") / O.point_cost(Xg, np.ones(d), False)
    O.first_argmax(acq)
    t_pts = time.perf_counter() - t0
    per_pt = t_pts / len(flat)
    out = {"points_per_s": G / (t_once + per_pt * G), "one_off_s": t_once, "per_point_s": per_pt, "sample_points": int(len(flat)),
           "threads": int(blas_threads), "cores": os.cpu_count()}
    if direct_pts > 0:
        # reference-faithful loop form (DoCalculus.py:50-89): one candidate at a time, mean and variance closures
        # evaluated separately.  Needs the Cholesky factor of the observational Gram.
        t0 = time.perf_counter()
        gp_full = O.obs_gp_fit(X, np.zeros(N), problem.s2, ls, problem.noise, form="diff", want_inverse=False)
        gp_full["alpha"] = problem.alpha_obs
        t_fit = time.perf_counter() - t0
        t0 = time.perf_counter()
        for x in Xg[:direct_pts]:
            O.do_prior_direct(gp_full, X, cols, x[None, :])   # mean closure
            O.do_prior_direct(gp_full, X, cols, x[None, :])   # variance closure (the reference runs predict twice)
        t_dir = (time.perf_counter() - t0) / direct_pts
        out["direct_form"] = {"points_per_s": 1.0 / t_dir, "candidates_timed": direct_pts, "obs_gp_factorisation_s": t_fit}
    return out


def run_reference(args, world, rank):
    if rank != 0:
        return
    from cbo_with_oop_b200.synthetic import scaled_set
    cfg, S = workload_config(args, world)
    dev = None   # the inputs' observational-GP state comes from host LAPACK here: none of this repository's kernels runs in this arm
    pr = scaled_set(0, n_obs=args.n_obs, p=args.p, d=D_INT, c=C_COND, n_int=N_INT, device=dev)
    sample = 2048 if args.n_obs >= 5000 else 16384
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = cpu_port(pr, sample, seed=i)
        if i >= args.warmup:
            vals.append(r["points_per_s"])
            times.append(time.perf_counter() - t0)
    v = float(np.mean(vals))
    sample_txt = ("each step: one-off precompute of 1 exploration set + %d seeded candidates of its grid, "
                  "extrapolated linearly to the set's %d candidates; all sets have the same cost" % (r["sample_points"], args.p ** D_INT))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample_txt},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of the reference's arithmetic (oracle/cbo_oracle.py, factorised vectorised form) on the "
                    "host cores; the reference itself needs GPy/emukit/paramz, which cannot be installed offline"}
    emit(line)


# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for ln in self.proc.stdout:
                self.rows.append([c.strip() for c in ln.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)]
        return {"sm_mhz": float(np.median(busy or sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# stdout carries exactly ONE line, the JSON record: the process-wide fd 1 is pointed at stderr before any library loads
# (NCCL prints its version banner to fd 1 from C), and the record goes to a private duplicate of the original stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def run_ours(args, world, rank, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from cbo_with_oop_b200.engine import SetProblem, SweepEngine
    from cbo_with_oop_b200.partition import SetSize, partition
    from cbo_with_oop_b200.synthetic import scaled_set

    cfg, S = workload_config(args, world)
    G_set = args.p ** D_INT
    sizes = [SetSize(G_set, args.n_obs, N_INT)] * S
    mine = partition(sizes, world)[rank]
    t_setup = time.time()
    problems = []
    for s in range(S):
        if mine[s][1] > 0:
            problems.append(scaled_set(s, n_obs=args.n_obs, p=args.p, d=D_INT, c=C_COND, n_int=N_INT, device=dev))
        else:  # shape-only placeholder: this rank never touches the set
            z = lambda *sh: np.broadcast_to(np.zeros(1), sh)
            problems.append(SetProblem(z(args.n_obs, D_INT), z(args.n_obs, C_COND), z(args.n_obs, C_COND), z(args.n_obs), z(1, 1),
                                       np.ones(D_INT), np.ones(C_COND), 1.0, [np.linspace(-2, 2, args.p)] * D_INT,
                                       z(N_INT, D_INT), z(N_INT), cost_fix=float(D_INT)))
    # the incumbent every rank would hold: min over the interventional outputs of ALL sets (cheap to regenerate)
    best = float(min(np.min(p.y_int) for p in problems if p.kyinv.shape[0] > 1))
    if world > 1:
        t = torch.tensor([best], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        best = float(t.item())
    eng = SweepEngine(problems, device=dev, rank=rank, world_size=world, pinned_staging=True)
    eng.timing = True
    t_setup = time.time() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(nsteps, with_upload):
        stage_ms, launches, h2d = {}, 0, 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
        out = None
        for _ in range(nsteps):
            if with_upload:
                h2d = eng.upload()
            l0 = eng.launches
            out = eng.sweep(best, "min")
            launches = eng.launches - l0
            for k, v in out.stage_ms.items():
                stage_ms[k] = stage_ms.get(k, 0.0) + v
        e1.record(torch.cuda.current_stream(dev))
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out, {k: v / nsteps for k, v in stage_ms.items()}, launches, h2d

    for _ in range(args.warmup):
        eng.sweep(best, "min")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, out, stage_ms, launches, _ = timed(args.steps, False)
    clocks = sampler.stop() if rank == 0 else None
    total_pts = S * G_set
    value = total_pts / (ms / args.steps * 1e-3)

    # secondary metric: the post-intervention trial (prior on the grid cached; one set gets a new interventional row and
    # is refitted, every other set only refreshes EI from its cached posterior).  Reported apart from the headline.
    refresh = None
    if eng.active:
        g0 = eng.active[0]
        pr0 = eng.problems[g0]
        x_old, y_old = pr0.x_int.copy(), pr0.y_int.copy()
        rng = np.random.default_rng(7)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nref = 5
        x_cur, y_cur = x_old, y_old
        for it in range(nref + 2):          # every trial appends one interventional row to the set, as CBO.intervene does
            if it == 2:
                r0.record(torch.cuda.current_stream(dev))
            x_cur = np.vstack([x_cur, rng.uniform(-2, 2, (1, D_INT))])
            y_cur = np.append(y_cur, 0.0)
            eng.set_interventional(g0, x_cur, y_cur)
            out_r = eng.refresh(best, "min", refit=[g0])
        r1.record(torch.cuda.current_stream(dev))
        barrier()
        eng.set_interventional(g0, x_old, y_old)
        ms_r = torch.tensor([r0.elapsed_time(r1) / nref], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_r, op=dist.ReduceOp.MAX)
        refresh = {"ms_per_trial": float(ms_r.item()), "value": total_pts / (float(ms_r.item()) * 1e-3), "unit": UNIT,
                   "stage_ms": out_r.stage_ms,
                   "what": "post-intervention trial: one interventional row appended to a set -> its interventional table, the prior "
                           "of the NEW row (u^T M u streams M once: HBM-bound), refit of that set, full posterior + EI for it, EI "
                           "refresh from cached mu/var (16 B/candidate) for the other sets, argmax"}

    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    ms_e, out_e, _, _, h2d = timed(e2e_steps, True)
    e2e_value = total_pts / (ms_e / e2e_steps * 1e-3)
    if world > 1:
        hb = torch.tensor([float(h2d)], dtype=torch.float64, device=dev)
        dist.all_reduce(hb, op=dist.ReduceOp.SUM)
        h2d = int(hb.item())
    d2h = world * eng.d2h_bytes_per_sweep

    line = None
    if rank == 0:
        # roofline of the dominant kernel (prior_eval on the grid): algorithmic flops per launch / live duration
        my_pts = sum(c for _, c in mine)
        N, d = args.n_obs, D_INT
        flops_alg = my_pts * (2.0 * N * N + 2.0 * N + d * N)          # SURVEY.md §8(d) F_prior, dense-counted
        nJ = (N + 127) // 128
        # what the symmetric kernel executes: the lower block triangle of M in 128 x 128 blocks; the last column block is
        # re-tiled over 16 / 32 / 64 columns when that covers its live columns (prior_eval.cu, consume_ragged_block)
        lc = N - (nJ - 1) * 128
        last_cols = 128 if (nJ == 1 or lc > 64) else (16 if lc <= 16 else 32 if lc <= 32 else 64)
        flops_exec = my_pts * 2.0 * (128 * 128 * (nJ - 1) * nJ / 2 + nJ * 128 * last_cols)
        dur_s = stage_ms["prior_eval_grid"] * 1e-3
        peak, peak_src = 36.97, "fallback constant"
        try:
            pk = json.load(open(FP64_PEAK_FILE))
            peak = max(r["tflops"] for r in pk["issue_rate"]["results"] if r["kind"].startswith("dmma"))
            peak_src = "measured FP64 DMMA issue rate on this pool's B200 (tools/fp64_peak.cu -> profiles/fp64_peak_r01.json; cuBLAS DGEMM reaches %.1f)" % pk["cublas_dgemm_burst_tflops"]
        except Exception:
            pass
        traffic = None
        try:
            traffic = json.load(open(TRAFFIC_FILE)).get("dram_bytes_per_launch")
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "cbo::prior_eval_kernel (which=0, grid)", "achieved": flops_alg / dur_s * 1e-12,
                "peak": peak, "unit": "TFLOP/s", "frac": flops_alg / dur_s * 1e-12 / peak, "traffic": traffic,
                "executed": flops_exec / dur_s * 1e-12, "executed_frac": flops_exec / dur_s * 1e-12 / peak,
                "kernel_ms_per_launch": stage_ms["prior_eval_grid"], "share_of_step": stage_ms["prior_eval_grid"] / (ms / args.steps),
                "peak_source": peak_src,
                "note": "FP64 tensor pipe (DMMA.8x8x4), not the bf16 figure of MEASURED_PEAKS.json. `achieved`/`frac` use the "
                        "dense-counted algorithmic figure of SURVEY.md 8(d), 2N^2+2N+dN flops per candidate; the kernel exploits the "
                        "symmetry of M and executes about N^2 of them, so `frac` can approach 2 -- `executed_frac` is the hardware "
                        "utilisation of the FP64 pipe."}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps, "api": "SweepEngine.upload() (pinned host -> device of every input incl. the N x N Ky^-1) "
                                                   "+ SweepEngine.sweep() -> host result"},
                "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
                "roofline": roof, "clocks": clocks, "stage_ms_per_step": stage_ms, "post_intervention_trial": refresh,
                "selected": {"set": out.set, "index": out.index, "value": out.value}, "setup_s": round(t_setup, 1)}
        if not args.no_cpu_baseline and world == 1:
            first = next(p for p in problems if p.kyinv.shape[0] > 1)
            r = cpu_port(first, 2048 if args.n_obs >= 5000 else 16384, direct_pts=1 if args.n_obs <= 10_000 else 0)
            line["cpu_baseline"] = {
                "value": r["points_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                "sample": "one-off precompute of 1 exploration set + %d seeded candidates of its grid with the vectorised factorised "
                          "NumPy/BLAS port (oracle/cbo_oracle.py), extrapolated linearly to the set's %d candidates"
                          % (r["sample_points"], G_set),
                "one_off_s": r["one_off_s"], "per_point_s": r["per_point_s"], "host_cores": r["cores"],
                "reference_faithful_direct_form": r.get("direct_form")}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if args.impl == "reference":
        run_reference(args, world, rank)
    else:
        run_ours(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
