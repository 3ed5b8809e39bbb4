"""GPU parity of the device-side observational-GP fit (cbo_obs_gp_fit: Gram, blocked Cholesky, triangular inverse,
Ky^-1 = L^-T L^-1, alpha) against the oracle's restatement of GPy's exact inference (oracle.obs_gp_fit, utils.py:40-45).
alpha and Ky^-1 are solutions of a system with condition number ~ N s2 / noise, so they are held to backward errors
(residuals) at n * eps level and to the 1e-6 forward rule; the downstream sweep must not notice which fit produced them."""
import numpy as np
import pytest

from helpers import RTOL, make_case, oracle_sweep, rel_err
from oracle import cbo_oracle as O

pytestmark = pytest.mark.gpu


def _data(seed, N, D, ard=True):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    y = np.sin(X @ rng.uniform(-1, 1, D)) + 0.1 * rng.standard_normal(N)
    s2 = float(rng.uniform(0.6, 1.6))
    ls = rng.uniform(0.7, 1.8, D) if ard else np.array([float(rng.uniform(0.8, 1.5))])
    return X, y, s2, ls


@pytest.mark.parametrize("N,D,ard", [(1, 1, False), (5, 2, True), (127, 3, True), (128, 1, False), (129, 6, True),
                                     (300, 4, True), (641, 12, True), (1000, 6, False)])
def test_fit_matches_oracle(cuda_engine_ready, N, D, ard):
    from cbo_with_oop_b200.obs_gp import fit_state_device
    X, y, s2, ls = _data(1000 + N, N, D, ard)
    ref = O.obs_gp_fit(X, y, s2, ls, form="diff")
    alpha, kyinv, tries = fit_state_device(X, y, s2, ls, 1e-2)
    alpha, kyinv = alpha.cpu().numpy(), kyinv.cpu().numpy()
    assert tries == ref["tries"] == 0
    Ky = O.rbf_K(X, X, s2, ls, "diff", same=True) + (1e-2 + 1e-8) * np.eye(N)
    nrm = np.abs(Ky).sum(1).max()
    # backward errors
    assert np.abs(Ky @ alpha - y).max() <= 1e-12 * (nrm * np.abs(alpha).max() + np.abs(y).max())
    assert np.abs(Ky @ kyinv - np.eye(N)).max() <= 1e-11 * nrm * np.abs(kyinv).max()
    np.testing.assert_array_equal(kyinv, kyinv.T)
    # forward errors against the oracle (cond ~ N s2 / noise <= 1e6 here)
    assert rel_err(alpha, ref["alpha"], 1e-6 * np.abs(ref["alpha"]).max()).max() <= RTOL
    assert rel_err(kyinv, ref["Kyinv"], 1e-6 * np.abs(ref["Kyinv"]).max()).max() <= RTOL


def test_duplicate_rows_need_no_jitter(cuda_engine_ready):
    """The reference's observe() appends the same rows again: exact duplicates in the design."""
    from cbo_with_oop_b200.obs_gp import fit_state_device
    X, y, s2, ls = _data(7, 150, 3)
    X = np.vstack([X, X[:40]]); y = np.concatenate([y, y[:40]])
    ref = O.obs_gp_fit(X, y, s2, ls, form="diff")
    alpha, kyinv, tries = fit_state_device(X, y, s2, ls, 1e-2)
    assert tries == ref["tries"]
    assert rel_err(alpha.cpu().numpy(), ref["alpha"], 1e-6 * np.abs(ref["alpha"]).max()).max() <= RTOL


def test_jitter_retry_rule(cuda_engine_ready):
    """noise = 0, 70 exactly duplicated rows and a variance of 1e12: the 1e-8 on the diagonal is far below the rounding of
    the pivots (1e12 * eps), so the plain factorisation meets a non-positive pivot; GPy's first jitter (mean(diag) * 1e-6)
    repairs it.  The device must walk the same retry ladder as jitchol."""
    from cbo_with_oop_b200.obs_gp import fit_state_device
    X, y, _, ls = _data(8, 140, 2)
    X[70:] = X[:70]
    ref = O.obs_gp_fit(X, y, 1e12, ls, noise=0.0, form="diff")
    _, _, tries = fit_state_device(X, y, 1e12, ls, 0.0)
    assert tries == ref["tries"] and tries >= 1


def test_sweep_with_device_fit_equals_sweep_with_supplied_state(cuda_engine_ready):
    from cbo_with_oop_b200.engine import SetProblem, SweepEngine
    kw, ora = make_case(seed=77, N=333, d=2, c=2, n=9, p=(17, 13))
    best = float(np.min(kw["y_int"]))
    ref = oracle_sweep(ora, best, "min")
    kw2 = dict(kw, alpha_obs=None, kyinv=None, y_obs=ora["gp"]["y"])
    eng = SweepEngine([SetProblem(**kw2)], keep=("acq",))
    out = eng.sweep(best, "min")
    assert out.index == ref["idx"]
    np.testing.assert_allclose(out.value, ref["val"], rtol=RTOL)
    assert rel_err(eng.fetch("m", 0), ref["mg"], 1e-6).max() <= RTOL
    assert rel_err(eng.fetch("v", 0), ref["vg"], 1e-6).max() <= RTOL


def test_large_fit_properties(cuda_engine_ready):
    """N = 4000 (32 panels): residuals of alpha and of sampled columns of Ky^-1, symmetry, agreement with torch's
    LAPACK-backed solve on the device (a second implementation, used here as a checker only)."""
    import torch
    from cbo_with_oop_b200.obs_gp import fit_state_device
    N, D = 4000, 6
    X, y, s2, ls = _data(9, N, D)
    alpha, kyinv, tries = fit_state_device(X, y, s2, ls, 1e-2)
    assert tries == 0
    Z = torch.as_tensor(X / ls, device="cuda:0")
    Ky = torch.cdist(Z, Z, compute_mode="donot_use_mm_for_euclid_dist").square_().mul_(-0.5).exp_().mul_(s2)
    Ky.diagonal().add_(1e-2 + 1e-8)
    yt = torch.as_tensor(y, device="cuda:0")
    nrm = float(Ky.abs().sum(1).max())
    assert float((Ky @ alpha - yt).abs().max()) <= 1e-12 * (nrm * float(alpha.abs().max()) + float(yt.abs().max()))
    cols = torch.arange(0, N, 97, device="cuda:0")
    R = Ky @ kyinv[:, cols]
    R[cols, torch.arange(len(cols), device="cuda:0")] -= 1.0
    assert float(R.abs().max()) <= 1e-11 * nrm * float(kyinv.abs().max())
    assert torch.equal(kyinv, kyinv.T)
    a2 = torch.cholesky_solve(yt[:, None], torch.linalg.cholesky(Ky))[:, 0]
    assert float(((alpha - a2).abs() / a2.abs().max()).max()) <= 1e-6


@pytest.mark.parametrize("N,D,ard", [(60, 1, False), (200, 3, True), (333, 6, False), (500, 12, True)])
def test_marginal_likelihood_and_gradient(cuda_engine_ready, N, D, ard):
    """cbo_obs_gp_nll against the host objective (GPy's formulas: obs_gp.neg_log_marginal_likelihood) and against a
    central finite difference of the device objective itself."""
    from cbo_with_oop_b200.obs_gp import DeviceObsGP, _device_objective, neg_log_marginal_likelihood
    X, y, s2, ls = _data(2000 + N, N, D, ard)
    theta = np.concatenate([[np.log(s2)], np.log(ls)])
    ref, gref = neg_log_marginal_likelihood(theta, X, y, ard, 1e-2)
    f = _device_objective(DeviceObsGP(X, y, 1e-2), ard)
    nll, g = f(theta)
    np.testing.assert_allclose(nll, ref, rtol=1e-9)
    np.testing.assert_allclose(g, gref, rtol=1e-6, atol=1e-8 * np.abs(gref).max())
    h = 1e-5
    for k in range(len(theta)):
        e = np.zeros_like(theta); e[k] = h
        fd = (f(theta + e)[0] - f(theta - e)[0]) / (2 * h)
        assert abs(fd - g[k]) <= 1e-5 * max(1.0, np.abs(g).max()), (k, fd, g[k])


def test_hyperparameter_search_on_device_matches_host(cuda_engine_ready):
    from cbo_with_oop_b200.obs_gp import optimize_hyperparameters
    X, y, _, _ = _data(31, 250, 2)
    floor = 0.5 * X.std(0)
    host = optimize_hyperparameters(X, y, 1.0, 1.0, True, 1e-2, min_lengthscale=floor, max_variance=20.0)
    dev = optimize_hyperparameters(X, y, 1.0, 1.0, True, 1e-2, min_lengthscale=floor, max_variance=20.0, device="cuda:0")
    np.testing.assert_allclose(dev[0], host[0], rtol=1e-3)
    np.testing.assert_allclose(dev[1], host[1], rtol=1e-3)
