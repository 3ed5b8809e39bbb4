"""Stage-by-stage timing of one rank's sweep on a config-5-shaped problem (development aid; bench.py is the
contract).  Example: python tools/perf_probe.py --n-obs 10000 --p 37 64 64 --sets 1 --variant 0"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

ap = argparse.ArgumentParser()
ap.add_argument("--n-obs", type=int, default=10000)
ap.add_argument("--p", type=int, nargs="+", default=[37, 64, 64])
ap.add_argument("--sets", type=int, default=1)
ap.add_argument("--n-int", type=int, default=32)
ap.add_argument("--c", type=int, default=3)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

import numpy as np
import torch

from cbo_with_oop_b200.engine import SweepEngine
from cbo_with_oop_b200.synthetic import best_of, scaled_set

t0 = time.time()
probs = []
for i in range(args.sets):
    pr = scaled_set(i, n_obs=args.n_obs, p=max(args.p), d=len(args.p), c=args.c, n_int=args.n_int, device="cuda:0")
    pr.grid = [np.linspace(-2.0, 2.0, pk) for pk in args.p]
    probs.append(pr)
setup_s = time.time() - t0
eng = SweepEngine(probs)
eng.timing = True
best = best_of(probs)
outs = []
for r in range(args.reps):
    torch.cuda.synchronize()
    t1 = time.time()
    out = eng.sweep(best, "min")
    torch.cuda.synchronize()
    outs.append((time.time() - t1, out))
wall, out = min(outs, key=lambda t: t[0])
G = sum(p.g_total for p in probs)
N = args.n_obs
res = {"variant": args.variant, "n_obs": N, "p": args.p, "sets": args.sets, "G": G, "setup_s": round(setup_s, 2),
       "wall_ms": round(wall * 1e3, 2), "stage_ms": {k: round(v, 3) for k, v in out.stage_ms.items()},
       "points_per_s": G / wall, "selected": [out.set, out.index, out.value]}
ms = out.stage_ms["prior_eval_grid"]
res["prior_eval_dense_tflops"] = G * (2.0 * N * N + 2 * N + len(args.p) * N) / ms * 1e-9
nJ = (N + 127) // 128
res["prior_eval_executed_tflops"] = G * 2.0 * (128 * 128 * nJ * (nJ + 1) / 2) / ms * 1e-9
print(json.dumps(res))
