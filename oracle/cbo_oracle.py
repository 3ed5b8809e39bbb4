"""CPU oracle for the CBO acquisition sweep -- TEST INFRASTRUCTURE, NOT THE PRODUCT.

This module restates, in NumPy/SciPy float64, the arithmetic of ChampiB/CBO_with_OOP's per-trial
hot path (causal prior -> per-set GP posterior -> Expected Improvement / cost -> argmax).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import it; the product (``cbo_with_oop_b200`` and ``src``) never does.

PARITY STATUS.  The reference ships no tests, golden vectors or fixtures for this path, and its
linear algebra lives in un-vendored third-party packages that are not installable here: GPy~=1.10.0,
emukit~=0.4.10, paramz~=0.9.5 (reference requirements.txt:6-9).  What pins this oracle:
  * PINNED against the reference's own code: tests/golden/reference_run_{toy,complete}.npz were produced by running
    the reference's modules unmodified (DoCalculus.compute_do, CausalRBF.K/Kdiag, GaussianProcessFactory.create,
    CausalExpectedImprovement, Cost, find_current_global, CBO.select_next_intervention) on the shipped data
    (tests/golden/make_reference_golden.py); tests/test_reference_run.py holds this oracle to them (literal form 1e-8,
    default form 1e-6, integers exact).
  * UNPINNED: GPy's internals themselves.  In that run GPy/emukit/paramz were replaced by the stand-in under
    tests/golden/gpy_standin, so ``Stationary._scaled_dist``, ``ExactGaussianInference.inference``, ``jitchol``,
    ``PosteriorExact._raw_predict`` and ``Gaussian.predictive_values`` are restated from their published algorithms
    (here and, independently, in the stand-in); the plain-RBF GP is cross-checked against scikit-learn's
    GaussianProcessRegressor in tests/test_oracle.py so the restatement is not purely self-certified.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import scipy.stats

OBS_NOISE = 1e-2       # utils.py:43  gp.likelihood.variance.fix(1e-2)
POST_NOISE = 1e-10     # GaussianProcessFactory.py:58,73  noise_var=1e-10
GPY_JITTER = 1e-8      # GPy ExactGaussianInference: diag.add(Ky, variance + 1e-8)


# --------------------------------------------------------------------------------------------------
# GPy building blocks
# --------------------------------------------------------------------------------------------------
def scaled_sqdist(X, X2, lengthscale, form="expanded"):
    """r^2 between rows of X (a,D) and X2 (b,D), scaled by the lengthscale(s).

    form="expanded": GPy Stationary._unscaled_dist with X2 given: ||x||^2+||y||^2-2x.y, clipped at 0
      (what causal_kernels.py:55 ``self._scaled_dist(X, X2)`` executes; for ARD the columns are divided
      by the lengthscales first, otherwise the distance is divided afterwards).
    form="diff": the same quantity from coordinate differences (no cancellation); used to bound how much
      of a GPU-vs-oracle gap is the reference's own rounding (SURVEY.md §7 "large-magnitude coordinates").
    """
    X = np.asarray(X, dtype=np.float64)
    X2 = np.asarray(X2, dtype=np.float64)
    ls = np.asarray(lengthscale, dtype=np.float64).reshape(-1)
    ard = ls.size > 1
    if form == "diff":
        l = ls if ard else ls[0]
        d = (X[:, None, :] - X2[None, :, :]) / l
        return np.sum(d * d, axis=2)
    if ard:
        X = X / ls
        X2 = X2 / ls
    x1sq = np.sum(np.square(X), 1)
    x2sq = np.sum(np.square(X2), 1)
    r2 = -2.0 * np.dot(X, X2.T) + (x1sq[:, None] + x2sq[None, :])
    r2 = np.clip(r2, 0, np.inf)
    if not ard:
        # GPy: sqrt(r2)/l then squared again in K_of_r; (sqrt(r2)/l)**2 is restated literally.
        r = np.sqrt(r2) / ls[0]
        return r * r
    r = np.sqrt(r2)
    return r * r


def rbf_K(X, X2, variance, lengthscale, form="expanded", same=False):
    """GPy RBF.K: variance*exp(-0.5 r^2) (causal_kernels.py:56,82).  same=True mimics X2=None
    (GPy zeroes the diagonal of r^2 in that branch)."""
    r2 = scaled_sqdist(X, X2, lengthscale, form)
    if same:
        np.fill_diagonal(r2, 0.0)
    return variance * np.exp(-0.5 * r2)


def jitchol(A, maxtries=5):
    """GPy util.linalg.jitchol: dpotrf; on failure add mean(diag)*1e-6 * 10^t jitter, t=0..maxtries-1.
    Returns (L lower, number of jitter retries that were needed)."""
    A = np.ascontiguousarray(A)
    try:
        return np.linalg.cholesky(A), 0
    except np.linalg.LinAlgError:
        pass
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    for t in range(1, maxtries + 1):
        try:
            return np.linalg.cholesky(A + np.eye(A.shape[0]) * jitter), t
        except np.linalg.LinAlgError:
            jitter *= 10
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


# --------------------------------------------------------------------------------------------------
# Observational GP (utils.py:40-45; GPy GPRegression exact inference), hyper-parameters are INPUTS
# --------------------------------------------------------------------------------------------------
def obs_gp_fit(X, y, variance, lengthscale, noise=OBS_NOISE, form="expanded", want_inverse=True):
    """Exact GP regression state for the observational GP of one exploration set.
    Ky = K + (noise + 1e-8) I ; L = jitchol(Ky) ; alpha = Ky^-1 y ; Kyinv = Ky^-1 (GPy pdinv/dpotri)."""
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    K = rbf_K(X, X, variance, lengthscale, form, same=True)
    Ky = K + (noise + GPY_JITTER) * np.eye(X.shape[0])
    L, tries = jitchol(Ky)
    alpha = scipy.linalg.cho_solve((L, True), y)
    out = dict(X=X, y=y, variance=float(variance), lengthscale=np.asarray(lengthscale, np.float64).reshape(-1),
               noise=float(noise), L=L, alpha=alpha, tries=tries, form=form)
    if want_inverse:
        Kyinv = scipy.linalg.cho_solve((L, True), np.eye(X.shape[0]))
        out["Kyinv"] = 0.5 * (Kyinv + Kyinv.T)
    return out


def obs_gp_predict(gp, Xnew):
    """GPy GP.predict (include_likelihood=True): mu = Kx^T alpha ; var = Kdiag - sum((L^-1 Kx)^2) + noise.
    Called by DoCalculus.py:77 ``gp.predict(intervened_inputs)``."""
    Kx = rbf_K(gp["X"], Xnew, gp["variance"], gp["lengthscale"], gp["form"])
    mu = Kx.T @ gp["alpha"]
    tmp = scipy.linalg.solve_triangular(gp["L"], Kx, lower=True)
    var = gp["variance"] - np.sum(np.square(tmp), 0) + gp["noise"]
    return mu, var


# --------------------------------------------------------------------------------------------------
# Causal prior (DoCalculus.py:34-89)
# --------------------------------------------------------------------------------------------------
def do_prior_direct(gp, cond_samples, intervened_cols, values):
    """Reference-faithful loop form.  For every row ``value`` of ``values`` (m,d): build the (S_mc, D)
    design whose intervened columns are the constant value[k] and whose other columns are the
    conditioning samples (DoCalculus.py:68-89, get_intervened_inputs), predict with the observational
    GP and average mean and variance over the samples (DoCalculus.py:59-60; the variance closure and the
    mean closure of the reference each run this and keep one half).

    cond_samples: (S_mc, D) array; only the non-intervened columns are read.
    Returns (m (len,), v (len,))."""
    values = np.atleast_2d(np.asarray(values, np.float64))
    S = cond_samples.shape[0]
    m = np.empty(values.shape[0])
    v = np.empty(values.shape[0])
    for i, value in enumerate(values):
        Z = np.array(cond_samples, dtype=np.float64, copy=True)
        for k, col in enumerate(intervened_cols):
            Z[:, col] = np.ones(S) * value[k]
        mu, var = obs_gp_predict(gp, Z)
        m[i] = np.mean(mu)
        v[i] = np.mean(var)
    return m, v


def prior_factors(gp, cond_samples, intervened_cols):
    """Factorised form of the same average.  With u_j(x)=prod_{k in I} exp(-.5 (x_k-X_jk)^2/l_k^2)
    and P_ji = prod_{k not in I} exp(-.5 (C_ik-X_jk)^2/l_k^2):
        w = s2 * alpha * mean_i P ;  M = s2^2 * Kyinv o (P P^T / S_mc)
        m(x) = u.w ;  v(x) = s2 + noise - u^T M u
    Returns dict(w, M, pbar)."""
    X = gp["X"]
    N, D = X.shape
    ls = gp["lengthscale"] if gp["lengthscale"].size > 1 else np.repeat(gp["lengthscale"], D)
    cond_cols = [c for c in range(D) if c not in intervened_cols]
    S = cond_samples.shape[0]
    if cond_cols:
        r2 = np.zeros((N, S))
        for c in cond_cols:
            d = (X[:, c][:, None] - cond_samples[:, c][None, :]) / ls[c]
            r2 += d * d
        P = np.exp(-0.5 * r2)
        pbar = P.mean(axis=1)
        Q = (P @ P.T) / S
    else:
        pbar = np.ones(N)
        Q = np.ones((N, N))
    s2 = gp["variance"]
    return dict(w=s2 * gp["alpha"] * pbar, M=(s2 * s2) * gp["Kyinv"] * Q, pbar=pbar)


def intervened_u(gp, intervened_cols, values):
    """u (m,N): u_gj = exp(-.5 sum_{k in I} ((x_gk - X_jk)/l_k)^2)."""
    X = gp["X"]
    D = X.shape[1]
    ls = gp["lengthscale"] if gp["lengthscale"].size > 1 else np.repeat(gp["lengthscale"], D)
    values = np.atleast_2d(np.asarray(values, np.float64))
    r2 = np.zeros((values.shape[0], X.shape[0]))
    for k, col in enumerate(intervened_cols):
        d = (values[:, k][:, None] - X[:, col][None, :]) / ls[col]
        r2 += d * d
    return np.exp(-0.5 * r2)


def do_prior_factorised(gp, factors, intervened_cols, values, chunk=4096, precise=False):
    """m, v at explicit points via the factorised form (best-effort vectorised CPU path).

    precise=True accumulates u.w and u^T M u in np.longdouble (64-bit significand on x86) and rounds once.  The quadratic
    form sums N^2 products of mixed sign; when the observational GP is confident, s2 + noise - u^T M u cancels to 1e-3 of
    its terms' total magnitude and more (coral data: condition number 4e8), so the float64 evaluation carries ~1e-8 of
    relative noise that the per-set fit then amplifies.  The reference evaluates the variance as a sum of squares
    (GPy: Kdiag - |L^-1 k|^2) and has no such noise; the golden fixtures therefore use precise=True wherever a value is
    stored (interventional rows, kept candidates)."""
    values = np.atleast_2d(np.asarray(values, np.float64))
    m = np.empty(values.shape[0])
    v = np.empty(values.shape[0])
    s2, noise = gp["variance"], gp["noise"]
    if precise:
        Ml, wl = factors["M"].astype(np.longdouble), factors["w"].astype(np.longdouble)
        for a in range(0, values.shape[0], 256):
            U = intervened_u(gp, intervened_cols, values[a:a + 256]).astype(np.longdouble)
            m[a:a + 256] = (U @ wl).astype(np.float64)
            q = np.einsum("gj,gj->g", U @ Ml, U)
            v[a:a + 256] = ((np.longdouble(s2) + np.longdouble(noise)) - q).astype(np.float64)
        return m, v
    for a in range(0, values.shape[0], chunk):
        U = intervened_u(gp, intervened_cols, values[a:a + chunk])
        m[a:a + chunk] = U @ factors["w"]
        v[a:a + chunk] = s2 + noise - np.einsum("gj,gj->g", U @ factors["M"], U)
    return m, v


# --------------------------------------------------------------------------------------------------
# Per-set GP on interventional data (GaussianProcessFactory.py:57-73, causal_kernels.py:45-79)
# --------------------------------------------------------------------------------------------------
def causal_K(X, X2, vX, vX2, form="expanded"):
    """CausalRBF.K (causal_kernels.py:45-62), lengthscale=1, variance=1 (GaussianProcessFactory.py:69):
    exp(-.5 r^2) + sqrt(v(X)) sqrt(v(X2))^T.  X2 is always passed explicitly (:53-55), so the diagonal of
    r^2 is not forced to zero."""
    K = rbf_K(X, X2, 1.0, 1.0, form)
    return K + np.outer(np.sqrt(vX), np.sqrt(vX2))


def posterior_fit(XI, yI, mI=None, vI=None, form="expanded"):
    """GPy exact inference for the per-set GP.  causal (mI, vI given): mean function m, CausalRBF kernel,
    noise 1e-10 (GaussianProcessFactory.py:63-73); non-causal (mI=vI=None): zero mean, RBF(l=1,s2=1),
    noise 1e-10 (:57-60; RBF.K(X) zeroes the diagonal of r^2).
    Ky = K + (1e-10 + 1e-8) I ; L = jitchol(Ky) ; alpha = Ky^-1 (y - m)."""
    XI = np.asarray(XI, np.float64)
    yI = np.asarray(yI, np.float64).reshape(-1)
    n = XI.shape[0]
    causal = mI is not None
    if causal:
        K = causal_K(XI, XI, vI, vI, form)
        resid = yI - mI
    else:
        K = rbf_K(XI, XI, 1.0, 1.0, form, same=True)
        resid = yI
    Ky = K + (POST_NOISE + GPY_JITTER) * np.eye(n)
    L, tries = jitchol(Ky)
    alpha = scipy.linalg.cho_solve((L, True), resid)
    return dict(XI=XI, L=L, alpha=alpha, tries=tries, vI=(np.asarray(vI, np.float64) if causal else None),
                causal=causal, form=form)


def posterior_predict(post, Xg, mg=None, vg=None):
    """GPy predict for the per-set GP (the call at causal_acquisition_functions.py:33):
    k* = K(X_I, x); mu = m(x) + k*^T alpha; var = Kdiag(x) - ||L^-1 k*||^2 + 1e-10, Kdiag = 1 + v(x)
    (causal_kernels.py:64-79); no clipping."""
    Xg = np.atleast_2d(np.asarray(Xg, np.float64))
    if post["causal"]:
        Ks = causal_K(post["XI"], Xg, post["vI"], vg, post["form"])
        kdiag = 1.0 + vg
        mean0 = mg
    else:
        Ks = rbf_K(post["XI"], Xg, 1.0, 1.0, post["form"])
        kdiag = np.ones(Xg.shape[0])
        mean0 = 0.0
    mu = Ks.T @ post["alpha"] + mean0
    tmp = scipy.linalg.solve_triangular(post["L"], Ks, lower=True)
    var = kdiag - np.sum(np.square(tmp), 0) + POST_NOISE
    return mu, var


# --------------------------------------------------------------------------------------------------
# Acquisition (causal_acquisition_functions.py:27-43,77-88; cost_functions.py:11-17; utils.py:29-37)
# --------------------------------------------------------------------------------------------------
def expected_improvement(mu, var, best, task="min"):
    """CausalExpectedImprovement.evaluate with jitter=0: sd=sqrt(var); u=(best-mu)/sd;
    EI = sd*(u*Phi(u)+phi(u)); negated when task != 'min' (:38-41)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        sd = np.sqrt(var)
        u = (best - mu) / sd
        pdf = scipy.stats.norm.pdf(u)
        cdf = scipy.stats.norm.cdf(u)
        ei = sd * (u * cdf + pdf)
    return ei if task == "min" else -ei


def point_cost(Xg, fix_costs, variable):
    """Cost per candidate: sum_k (fix_k + [variable] |x_k|)  (GraphInterface.py:46-50 applied per point;
    the reference sums |x| over the whole batch, Appendix B #12 -- the per-point form is the stated intent)."""
    Xg = np.atleast_2d(np.asarray(Xg, np.float64))
    c = float(np.sum(fix_costs)) * np.ones(Xg.shape[0])
    if variable:
        c = c + np.sum(np.abs(Xg), axis=1)
    return c


def first_argmax(a):
    """np.argmax semantics with NaN treated as -inf (SURVEY.md §7 "NaN policy").  Returns (index, value, n_nan)."""
    a = np.asarray(a, np.float64)
    nan = np.isnan(a)
    b = np.where(nan, -np.inf, a)
    i = int(np.argmax(b))
    return i, float(b[i]), int(nan.sum())


def tensor_grid(tables):
    """C-ordered tensor product of per-dimension coordinate tables, last dimension fastest
    (index = ((i0*p1)+i1)*p2+i2)."""
    mesh = np.meshgrid(*tables, indexing="ij")
    return np.stack([g.reshape(-1) for g in mesh], axis=1)


def sweep_set(gp, cond_samples, intervened_cols, XI, yI, tables, best, task="min", fix_costs=None,
              variable_cost=False, causal=True, prior="factorised", factors=None, form="expanded", precise_int=False):
    """One exploration set's share of a trial: prior on X_I and on the grid, posterior fit, predict,
    EI / cost and the within-set first argmax.  Returns a dict of every intermediate (for parity tests).
    precise_int: evaluate the prior at the interventional rows with do_prior_factorised(precise=True)."""
    Xg = tensor_grid(tables)
    d = Xg.shape[1]
    fix_costs = np.ones(d) if fix_costs is None else fix_costs
    out = {}
    if causal:
        if prior == "factorised":
            factors = prior_factors(gp, cond_samples, intervened_cols) if factors is None else factors
            mI, vI = do_prior_factorised(gp, factors, intervened_cols, XI, precise=precise_int)
            mg, vg = do_prior_factorised(gp, factors, intervened_cols, Xg)
        else:
            mI, vI = do_prior_direct(gp, cond_samples, intervened_cols, XI)
            mg, vg = do_prior_direct(gp, cond_samples, intervened_cols, Xg)
        post = posterior_fit(XI, yI, mI, vI, form)
        mu, var = posterior_predict(post, Xg, mg, vg)
        out.update(mI=mI, vI=vI, mg=mg, vg=vg)
    else:
        post = posterior_fit(XI, yI, form=form)
        mu, var = posterior_predict(post, Xg)
    ei = expected_improvement(mu, var, best, task)
    acq = ei / point_cost(Xg, fix_costs, variable_cost)
    idx, val, n_nan = first_argmax(acq)
    out.update(L=post["L"], alpha=post["alpha"], tries=post["tries"], mu=mu, var=var, ei=ei, acq=acq,
               idx=idx, val=val, n_nan=n_nan, x=Xg[idx], post=post, factors=factors if causal else None)
    return out


def select_set(values):
    """CBO.select_next_intervention (CBO.py:269-277): first set attaining the maximum; NaN = -inf."""
    i, v, _ = first_argmax(np.asarray(values, np.float64))
    return i, v
