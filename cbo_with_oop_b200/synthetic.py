"""Seeded synthetic sweeps of the shapes BASELINE.json names.  No dataset, no network."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .engine import SetProblem
from .obs_gp import fit_state


def scaled_set(index: int, n_obs: int = 10_000, p: int = 100, d: int = 3, c: int = 3, n_int: int = 32,
               seed: int = 1005, device=None, device_fit: bool = False) -> SetProblem:
    """One exploration set of config 5 ("synthetic scaled sweep"): X_obs ~ N(0,1)^(N x (d+c)),
    y = sin(x.a) + 0.1 eps, a ~ U(-1,1); s2 = 1, l = 1, noise 1e-2; grid [-2,2]^d with p points per dim;
    n_int distinct grid points as interventional data with y_I from the same function; cost type 1.
    device_fit: hand (X, y) to the engine, which forms alpha and Ky^-1 on the device (cbo_obs_gp_fit) -- no N x N host array."""
    rng = np.random.default_rng([seed, index])
    D = d + c
    X = rng.standard_normal((n_obs, D))
    a = rng.uniform(-1.0, 1.0, D)
    y = np.sin(X @ a) + 0.1 * rng.standard_normal(n_obs)
    alpha, kyinv = (None, None) if device_fit else fit_state(X, y, 1.0, np.ones(D), 1e-2, device=device)
    grid = [np.linspace(-2.0, 2.0, p) for _ in range(d)]
    flat = rng.choice(p ** d, size=n_int, replace=False)
    ii = np.stack(np.unravel_index(flat, (p,) * d), axis=1)
    x_int = np.stack([grid[k][ii[:, k]] for k in range(d)], axis=1)
    # E[y | do(x)] of the generating function, averaged over the conditioning columns' law
    zc = rng.standard_normal((256, c)) if c else np.zeros((1, 0))
    y_int = np.array([np.mean(np.sin(x @ a[:d] + zc @ a[d:])) for x in x_int]) + 0.01 * rng.standard_normal(n_int)
    return SetProblem(x_obs_int=X[:, :d], x_obs_cond=X[:, d:], mc_cond=X[:, d:], alpha_obs=alpha, kyinv=kyinv,
                      ls_int=np.ones(d), ls_cond=np.ones(c), s2=1.0, grid=grid, x_int=x_int, y_int=y_int,
                      cost_fix=float(d), name=f"synthetic{index}", y_obs=y if device_fit else None)


def scaled_sweep(num_sets: int = 16, first: int = 0, **kw) -> List[SetProblem]:
    return [scaled_set(first + i, **kw) for i in range(num_sets)]


def best_of(problems: List[SetProblem], task: str = "min") -> float:
    ys = np.concatenate([p.y_int for p in problems])
    return float(ys.min() if task == "min" else ys.max())
