"""Observational-GP state (alpha = Ky^-1 y, Ky^-1) for frozen hyper-parameters.

These are INPUTS of the sweep: the reference produces them with GPy's exact inference inside
fit_gaussian_process (utils.py:40-45) whenever the agent observes.  `fit_state(..., device=...)` runs the library's own
blocked Cholesky / triangular inverse on the FP64 tensor pipe (csrc/obs_gp_fit.cu, cbo_obs_gp_fit);
without a device the small shipped graphs go through host SciPy (host code producing inputs, never on the sweep's path).
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
        # host path: input preparation for fixtures / small shipped graphs (the agent fits on the device, CBO.observe).
        # Same jitter ladder as the device path and GPy's jitchol: mean(diag) * 1e-6 * 10^t, t = 0..4.
        Ky = rbf_gram(X, s2, ls) + (noise + GPY_JITTER) * np.eye(N)
        jitter, L = 0.0, None
        for t in range(6):
            try:
                L = np.linalg.cholesky(Ky + jitter * np.eye(N) if jitter else Ky)
                break
            except np.linalg.LinAlgError:
                if t == 5 or np.any(np.diag(Ky) <= 0.0):
                    raise np.linalg.LinAlgError("observational Gram matrix not positive definite, even with jitter")
                jitter = float(np.mean(np.diag(Ky))) * 1e-6 if t == 0 else jitter * 10.0
        kyinv = scipy.linalg.cho_solve((L, True), np.eye(N))
        kyinv = 0.5 * (kyinv + kyinv.T)
        return kyinv @ y, kyinv
    alpha, kyinv, _ = fit_state_device(X, y, s2, ls, noise, device)
    return alpha.cpu().numpy(), kyinv.cpu().numpy()


class DeviceObsGP:
    """Device-resident exact-inference state of one observational GP (cbo_obs_gp_fit / cbo_obs_gp_nll, csrc/obs_gp_fit.cu).
    Holds the design, the outputs and the workspace, so that a hyper-parameter search re-fits in place."""

    def __init__(self, X, y, noise: float = 1e-2, device="cuda:0"):
        import ctypes as C

        import torch

        from . import _lib
        self._C, self._torch, self._lib = C, torch, _lib
        self.lib = _lib.load()
        self.dev = torch.device(device)
        X = np.ascontiguousarray(np.asarray(X, np.float64))
        self.N, self.D = X.shape
        if self.D > _lib.CBO_MAX_D + _lib.CBO_MAX_C:
            raise ValueError(f"the observational GP has {self.D} input columns; the library takes {_lib.CBO_MAX_D + _lib.CBO_MAX_C}")
        N, d = self.N, min(self.D, _lib.CBO_MAX_D)
        self.d, self.noise = d, float(noise)
        self.h = (_lib.SetDesc * 1)()
        S = self.h[0]
        S.d, S.c, S.n_obs, S.n_obs_pad, S.n_int, S.causal = d, self.D - d, N, -(-N // _lib.CBO_NPAD) * _lib.CBO_NPAD, 1, 1
        S.n_mc, S.n_mc_pad = 1, _lib.CBO_SPAD
        S.g_total, S.g_begin, S.g_count = 1, 0, 0
        for k in range(_lib.CBO_MAX_D):
            S.p[k] = 1
        S.noise, S.cost_fix = self.noise, 1.0
        self.xt = torch.from_numpy(np.ascontiguousarray(X.T)).to(self.dev)     # (D, N): one contiguous row per input column
        self.yt = torch.from_numpy(np.ascontiguousarray(np.asarray(y, np.float64).reshape(-1))).to(self.dev)
        self.alpha = torch.empty((N,), dtype=torch.float64, device=self.dev)
        self.kyinv = torch.empty((N, N), dtype=torch.float64, device=self.dev)
        self.info = torch.zeros((1,), dtype=torch.int32, device=self.dev)
        self.out = torch.zeros((2 + self.D,), dtype=torch.float64, device=self.dev)
        S.x_obs_int, S.x_obs_cond = self.xt.data_ptr(), self.xt.data_ptr() + d * N * 8
        S.y_obs, S.alpha_obs, S.kyinv = self.yt.data_ptr(), self.alpha.data_ptr(), self.kyinv.data_ptr()
        self._set_hyper(1.0, np.ones(self.D))
        self.ws = torch.empty((self.lib.cbo_obs_gp_workspace_bytes(self.h, 1),), dtype=torch.uint8, device=self.dev)
        self.tries = 0

    def _set_hyper(self, s2, ls):
        S = self.h[0]
        ls = np.broadcast_to(np.asarray(ls, np.float64).reshape(-1), (self.D,))
        for k in range(self.d):
            S.ls_int[k] = float(ls[k])
        for k in range(self.D - self.d):
            S.ls_cond[k] = float(ls[self.d + k])
        S.s2 = float(s2)

    def _stream(self):
        return self._C.c_void_p(self._torch.cuda.current_stream(self.dev).cuda_stream)

    def fit(self, s2: float, ls, max_tries: int = 5):
        """alpha, Ky^-1 for the given hyper-parameters.  GPy's jitchol rule on a non-positive pivot: retry with
        mean(diag Ky) * 1e-6 * 10^t added to the diagonal, t = 0..4.  Returns the number of retries used."""
        C = self._C
        self._set_hyper(s2, ls)
        jitter, tries = 0.0, 0
        while True:
            self._lib.check(self.lib.cbo_obs_gp_fit(self.h, 1, jitter, C.c_void_p(self.ws.data_ptr()), self.ws.numel(),
                                                    C.c_void_p(self.info.data_ptr()), self._stream()), "cbo_obs_gp_fit")
            if int(self.info.item()) == 0:
                self.tries = tries
                return tries
            if tries == max_tries:
                raise np.linalg.LinAlgError("observational Gram matrix not positive definite, even with jitter")
            jitter = (s2 + self.noise + GPY_JITTER) * 1e-6 if tries == 0 else jitter * 10.0
            tries += 1

    def nll_grad(self):
        """(-log p(y | X, s2, l), d/dlog s2, d/dlog l_k (D,)) of the state the last fit() produced."""
        C = self._C
        self._lib.check(self.lib.cbo_obs_gp_nll(self.h, C.c_void_p(self.ws.data_ptr()), self.ws.numel(),
                                                C.c_void_p(self.out.data_ptr()), self._stream()), "cbo_obs_gp_nll")
        o = self.out.cpu().numpy()
        return float(o[0]), float(o[1]), o[2:2 + self.D].copy()


def fit_state_device(X, y, s2: float, ls, noise: float = 1e-2, device="cuda:0", max_tries: int = 5):
    """(alpha (N,), kyinv (N, N), jitter retries) as float64 torch tensors on `device`, computed by cbo_obs_gp_fit."""
    gp = DeviceObsGP(X, y, noise, device)
    tries = gp.fit(s2, ls, max_tries)
    return gp.alpha, gp.kyinv, tries


def neg_log_marginal_likelihood(theta, X, y, ard, noise):
    """-log p(y | X, s2, l) and its gradient w.r.t. theta = log(s2), log(l) (l scalar or per-dimension).
    Zero mean, RBF kernel, fixed Gaussian noise (utils.py:41-43)."""
    N, D = X.shape
    s2 = np.exp(theta[0])
    ls = np.exp(theta[1:]) if ard else np.repeat(np.exp(theta[1]), D)
    Z = X / ls
    sq = np.sum(Z * Z, 1)
    r2 = np.clip(sq[:, None] + sq[None, :] - 2.0 * (Z @ Z.T), 0.0, None)
    np.fill_diagonal(r2, 0.0)
    K = s2 * np.exp(-0.5 * r2)
    Ky = K + (noise + GPY_JITTER) * np.eye(N)
    try:
        L = np.linalg.cholesky(Ky)
    except np.linalg.LinAlgError:
        return 1e25, np.zeros_like(theta)
    alpha = scipy.linalg.cho_solve((L, True), y)
    nll = 0.5 * y @ alpha + np.sum(np.log(np.diag(L))) + 0.5 * N * np.log(2 * np.pi)
    Kinv = scipy.linalg.cho_solve((L, True), np.eye(N))
    W = np.outer(alpha, alpha) - Kinv            # dL/dK = 0.5 * W
    g = np.empty_like(theta)
    g[0] = -0.5 * np.sum(W * K)                  # dK/dlog s2 = K
    if ard:
        for k in range(D):
            dk = (Z[:, k][:, None] - Z[:, k][None, :]) ** 2   # dK/dlog l_k = K * dk
            g[1 + k] = -0.5 * np.sum(W * K * dk)
    else:
        g[1] = -0.5 * np.sum(W * K * r2)
    return nll, g


def _device_objective(gp: "DeviceObsGP", ard: bool):
    """theta = (log s2, log l...) -> (nll, gradient) evaluated on the device (one K5 fit per call)."""
    def f(theta):
        s2 = float(np.exp(theta[0]))
        ls = np.exp(theta[1:]) if ard else np.repeat(np.exp(theta[1]), gp.D)
        try:
            gp.fit(s2, ls, max_tries=0)          # like the host objective: a failed factorisation is a rejected point
        except np.linalg.LinAlgError:
            return 1e25, np.zeros_like(theta)
        nll, g_s2, g_l = gp.nll_grad()
        return nll, np.concatenate([[g_s2], g_l if ard else [g_l.sum()]])
    return f


def optimize_hyperparameters(X, y, s2=1.0, ls=1.0, ard=False, noise=1e-2, max_iters=1000, min_lengthscale=None,
                             max_variance=None, device=None):
    """Stand-in for GPy's `gp.optimize()` in fit_gaussian_process (utils.py:40-45): maximise the marginal
    likelihood over the RBF variance and lengthscale(s), Gaussian noise fixed.  L-BFGS-B in log space from the
    given start (one run, like GPy's default).  Host SciPy -- produces INPUTS of the sweep, not part of it."""
    import scipy.optimize
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    D = X.shape[1]
    ls0 = np.broadcast_to(np.asarray(ls, np.float64).reshape(-1), (D,)) if ard else np.asarray(ls, np.float64).reshape(-1)[:1]
    theta0 = np.concatenate([[np.log(s2)], np.log(ls0)])
    # GPy optimises without bounds; +-12 in log space only keeps the search finite.  `min_lengthscale` (scalar or per
    # dimension) is an optional floor: with the noise pinned to 1e-2 the likelihood of noisy data is maximised by a
    # vanishing lengthscale, which makes the causal prior constant -- callers that want a usable prior set a floor.
    bounds = [(-12.0, 12.0)] * theta0.size
    if min_lengthscale is not None:
        lo = np.log(np.broadcast_to(np.asarray(min_lengthscale, np.float64).reshape(-1), (theta0.size - 1,)))
        bounds = [bounds[0]] + [(float(l), 12.0) for l in lo]
        theta0[1:] = np.maximum(theta0[1:], lo)
    if max_variance is not None:
        bounds[0] = (-12.0, float(np.log(max_variance)))
    if device is not None:   # the objective and its gradient on the device (cbo_obs_gp_fit + cbo_obs_gp_nll); same search
        res = scipy.optimize.minimize(_device_objective(DeviceObsGP(X, y, noise, device), ard), theta0, jac=True,
                                      method="L-BFGS-B", bounds=bounds, options={"maxiter": max_iters})
    else:
        res = scipy.optimize.minimize(neg_log_marginal_likelihood, theta0, args=(X, y, ard, noise), jac=True, method="L-BFGS-B",
                                      bounds=bounds, options={"maxiter": max_iters})
    th = res.x if np.isfinite(res.fun) and res.fun < 1e24 else theta0
    s2_opt = float(np.exp(th[0]))
    ls_opt = np.exp(th[1:]) if ard else np.repeat(np.exp(th[1]), D)
    return s2_opt, ls_opt
