"""Exploration-set problems of the shipped-data configurations (BASELINE.json configs 1-4) from the committed fixtures
tests/golden/golden_*.npz (written by tests/golden/make_golden.py from the reference's shipped data): the inputs only --
observational design, frozen hyper-parameters, interventional rows, grid ranges.  Used by bench.py's `small_configs`
block and by tools/golden_probe.py; the expected outputs in the same files are read by the tests alone."""
from __future__ import annotations

import os
from typing import List

import numpy as np

from .engine import SetProblem

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CONFIGS = {"toy": "toy_graph (configs[0])", "simplified_coral": "simplified_coral_graph (configs[1])",
           "complete": "complete_graph (configs[2])", "coral_synth": "coral_graph, synthetic observations (configs[3])"}


def golden_path(config: str) -> str:
    return os.path.join(GOLDEN_DIR, f"golden_{config}.npz")


def load_golden_problems(config: str, device_fit: bool = True):
    """(problems, best, task).  device_fit: the observational state (alpha, Ky^-1) is produced on the device from (X, y)
    by cbo_obs_gp_fit whenever the fixture stores y_obs; otherwise the stored host arrays are used."""
    z = np.load(golden_path(config), allow_pickle=False)
    problems: List[SetProblem] = []
    for s in range(int(z["num_sets"])):
        k = f"set{s}_"
        state = {}
        if device_fit and k + "y_obs" in z:
            state = dict(alpha_obs=None, kyinv=None, y_obs=z[k + "y_obs"].reshape(-1))
        elif k + "kyinv" in z:
            state = dict(alpha_obs=z[k + "alpha_obs"], kyinv=z[k + "kyinv"])
        else:
            from .obs_gp import fit_state
            a, ki = fit_state(np.hstack([z[k + "x_obs_int"], z[k + "x_obs_cond"]]), z[k + "y_obs"], float(z[k + "s2"]),
                              np.concatenate([z[k + "ls_int"], z[k + "ls_cond"]]), 1e-2)
            state = dict(alpha_obs=a, kyinv=ki)
        grid = [np.linspace(lo, hi, int(p)) for lo, hi, p in z[k + "grid_lo_hi_p"]]
        problems.append(SetProblem(x_obs_int=z[k + "x_obs_int"], x_obs_cond=z[k + "x_obs_cond"], mc_cond=z[k + "x_obs_cond"],
                                   ls_int=z[k + "ls_int"], ls_cond=z[k + "ls_cond"], s2=float(z[k + "s2"]), grid=grid,
                                   x_int=z[k + "x_int"], y_int=z[k + "y_int"], cost_fix=float(z[k + "cost_fix"]),
                                   name=str(z[k + "name"]), **state))
    return problems, float(z["best"]), str(z["task"])
