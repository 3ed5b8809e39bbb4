"""Observational-GP state (alpha = Ky^-1 y, Ky^-1) for frozen hyper-parameters.

These are INPUTS of the sweep: the reference produces them with GPy's exact inference inside
fit_gaussian_process (utils.py:40-45) whenever the agent observes.  Host SciPy for the shipped graph sizes;
for N ~ 1e4 the O(N^3) factorisation runs through torch.linalg on the device (plumbing, outside the timed
sweep; listed as the next row to move into csrc/ in SURVEY.md §8f.2).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

GPY_JITTER = 1e-8


def rbf_gram(X: np.ndarray, s2: float, ls: np.ndarray) -> np.ndarray:
    Z = np.asarray(X, np.float64) / np.asarray(ls, np.float64).reshape(1, -1)
    sq = np.sum(Z * Z, 1)
    r2 = sq[:, None] + sq[None, :] - 2.0 * (Z @ Z.T)
    np.fill_diagonal(r2, 0.0)
    return s2 * np.exp(-0.5 * np.clip(r2, 0.0, None))


def fit_state(X: np.ndarray, y: np.ndarray, s2: float, ls, noise: float = 1e-2, device=None):
    """Returns (alpha (N,), kyinv (N,N)) as float64 NumPy arrays."""
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    N, D = X.shape
    ls = np.broadcast_to(np.asarray(ls, np.float64).reshape(-1), (D,)) if np.size(ls) in (1, D) else None
    if ls is None:
        raise ValueError("lengthscale must have 1 or D entries")
    if device is None:
        Ky = rbf_gram(X, s2, ls) + (noise + GPY_JITTER) * np.eye(N)
        L = np.linalg.cholesky(Ky)
        kyinv = scipy.linalg.cho_solve((L, True), np.eye(N))
        kyinv = 0.5 * (kyinv + kyinv.T)
        return kyinv @ y, kyinv
    import torch
    Z = torch.as_tensor(X / ls.reshape(1, -1), device=device)
    r2 = torch.cdist(Z, Z, compute_mode="donot_use_mm_for_euclid_dist").square_()
    Ky = r2.mul_(-0.5).exp_().mul_(s2)
    Ky.diagonal().add_(noise + GPY_JITTER)
    L = torch.linalg.cholesky(Ky)
    del Ky
    kyinv = torch.cholesky_inverse(L)
    del L
    kyinv = 0.5 * (kyinv + kyinv.T)
    alpha = kyinv @ torch.as_tensor(y, device=device)
    return alpha.cpu().numpy(), kyinv.cpu().numpy()
