"""Timing of the device-side observational-GP fit (cbo_obs_gp_fit, csrc/obs_gp_fit.cu) at the config-5 size, beside
torch.linalg (cuSOLVER) on the same matrix as a yardstick.  python tools/fit_probe.py --n-obs 10000"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
ap = argparse.ArgumentParser()
ap.add_argument("--n-obs", type=int, default=10000)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

import numpy as np
import torch

from cbo_with_oop_b200 import _lib
from cbo_with_oop_b200.obs_gp import DeviceObsGP, fit_state_device

N, D = args.n_obs, 6
rng = np.random.default_rng(5)
X = rng.standard_normal((N, D))
y = np.sin(X @ rng.uniform(-1, 1, D)) + 0.1 * rng.standard_normal(N)
lib = _lib.load()
res = {"n_obs": N, "ours_ms": [], "launches": [], "torch_potrf_potri_ms": []}
for _ in range(args.reps):
    l0 = lib.cbo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    alpha, kyinv, tries = fit_state_device(X, y, 1.0, np.ones(D), 1e-2)
    e1.record()
    torch.cuda.synchronize()
    res["ours_ms"].append(e0.elapsed_time(e1))
    res["launches"].append(int(lib.cbo_launch_count() - l0))
# the fit alone on a resident design (what a hyper-parameter search pays per evaluation): no allocation, no upload
gp = DeviceObsGP(X, y, 1e-2, "cuda:0")
res["panel_blocks"] = int(os.environ.get("CBO_FIT_PANEL", "4"))
res["ours_fit_only_ms"] = []
for _ in range(args.reps + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    gp.fit(1.0, np.ones(D))
    e1.record()
    torch.cuda.synchronize()
    res["ours_fit_only_ms"].append(e0.elapsed_time(e1))
res["fit_only_vs_whole_call_max_abs_diff"] = float((gp.kyinv - kyinv).abs().max())
del gp
Z = torch.as_tensor(X, device="cuda:0")
Ky = torch.cdist(Z, Z, compute_mode="donot_use_mm_for_euclid_dist").square_().mul_(-0.5).exp_()
Ky.diagonal().add_(1e-2 + 1e-8)
for _ in range(args.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    L = torch.linalg.cholesky(Ky)
    Ki = torch.cholesky_inverse(L)
    e1.record()
    torch.cuda.synchronize()
    res["torch_potrf_potri_ms"].append(e0.elapsed_time(e1))
res["max_abs_diff_vs_torch"] = float((Ki - kyinv).abs().max())
res["flops_dense_counted"] = N ** 3          # potrf N^3/3 + trtri N^3/3 + lauum N^3/3
res["ours_tflops"] = N ** 3 / (min(res["ours_fit_only_ms"]) * 1e-3) / 1e12
print(json.dumps(res))
