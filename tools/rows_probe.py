"""Short driver for profiling the interventional-row prior kernel (prior_rows.cu) at the config-5 observational size:
one full trial on a small grid (wide kernel: all 32 rows), then post-intervention trials that append one row each
(narrow kernel, one pass over M).  Example: python tools/rows_probe.py --n-obs 10000 --appends 3"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

ap = argparse.ArgumentParser()
ap.add_argument("--n-obs", type=int, default=10000)
ap.add_argument("--p", type=int, default=20)
ap.add_argument("--appends", type=int, default=3)
args = ap.parse_args()

import numpy as np
import torch

from cbo_with_oop_b200.engine import SweepEngine
from cbo_with_oop_b200.synthetic import scaled_set

pr = scaled_set(0, n_obs=args.n_obs, p=args.p, d=3, c=3, n_int=32, device="cuda:0")
eng = SweepEngine([pr])
eng.timing = True
best = float(pr.y_int.min())
res = {"n_obs": args.n_obs, "full_trial_stage_ms": [], "append_stage_ms": []}
for _ in range(3):
    res["full_trial_stage_ms"].append(eng.sweep(best, "min").stage_ms)
rng = np.random.default_rng(3)
x, y = pr.x_int.copy(), pr.y_int.copy()
for _ in range(args.appends):
    x = np.vstack([x, rng.uniform(-2, 2, (1, 3))])
    y = np.append(y, 0.0)
    eng.set_interventional(0, x, y)
    res["append_stage_ms"].append(eng.refresh(best, "min", refit=[0]).stage_ms)
torch.cuda.synchronize()
N = args.n_obs
npad = -(-N // 128) * 128
nJ = npad // 128
tri_bytes = 8 * 128 * 128 * nJ * (nJ + 1) // 2
res["M_lower_block_triangle_bytes"] = tri_bytes
res["append_prior_GBps"] = [tri_bytes / (s["prior_eval_train"] * 1e-3) / 1e9 for s in res["append_stage_ms"]]
print(json.dumps(res))
