"""Stage timings of one full trial on the shipped-data golden configurations (BASELINE.json configs 1-4).
python tools/golden_probe.py [toy complete simplified_coral coral_synth]"""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from test_golden_gpu import GOLD, load_problems
from cbo_with_oop_b200.engine import SweepEngine

for config in (sys.argv[1:] or ["toy", "complete", "simplified_coral", "coral_synth"]):
    z = np.load(os.path.join(GOLD, f"golden_{config}.npz"), allow_pickle=False)
    problems = load_problems(z)
    eng = SweepEngine(problems)
    eng.timing = True
    best = float(z["best"])
    res = []
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.sweep(best, "min")
        e1.record()
        torch.cuda.synchronize()
        res.append((e0.elapsed_time(e1), out.stage_ms))
    ms, st = min(res, key=lambda t: t[0])
    G = sum(p.g_total for p in problems)
    N = max(p.x_obs_int.shape[0] for p in problems)
    flops = sum(p.g_total * p.x_obs_int.shape[0] ** 2 for p in problems)
    print(json.dumps({"config": config, "sets": len(problems), "candidates": G, "n_obs_max": N, "ms_per_trial": round(ms, 3),
                      "grid_points_per_s": G / (ms * 1e-3), "executed_tflops_prior": flops / (st["prior_eval_grid"] * 1e-3) / 1e12,
                      "stage_ms": {k: round(v, 3) for k, v in st.items()}}))
