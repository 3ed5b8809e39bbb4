"""CPU tests of the oracle (oracle/cbo_oracle.py): internal consistency (reference-faithful loop form vs factorised
form), an independent second source for the plain-RBF GP (scikit-learn), closed-form / SciPy checks of EI, argmax
semantics, and reproduction of the committed golden vectors.  The reference ships no known-answer tests for this path
(SURVEY.md §4); the statistical anchor it does ship -- the true causal effect of do(Z) on the toy graph -- is used
as a loose sanity check of the causal prior."""
import os

import numpy as np
import pytest
import scipy.stats

from helpers import make_case, oracle_sweep
from oracle import cbo_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_direct_and_factorised_prior_agree():
    for seed, d, c, S in [(1, 1, 0, None), (2, 2, 2, None), (3, 3, 1, 37), (4, 1, 3, 500)]:
        kw, ora = make_case(seed, N=80, d=d, c=c, n=5, p=(4,) * d, S_mc=S)
        gp, cond, cols = ora["gp"], ora["cond"], ora["cols"]
        vals = np.random.default_rng(seed).uniform(-2, 2, (6, d))
        m1, v1 = O.do_prior_direct(gp, cond, cols, vals)
        m2, v2 = O.do_prior_factorised(gp, O.prior_factors(gp, cond, cols), cols, vals)
        np.testing.assert_allclose(m1, m2, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(v1, v2, rtol=1e-9, atol=1e-12)
        assert np.all(v1 >= gp["noise"] * (1 - 1e-9))     # a mean of predictive variances incl. noise


def test_plain_rbf_gp_against_sklearn():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel
    rng = np.random.default_rng(0)
    X = rng.normal(size=(150, 3)); y = np.sin(X @ [0.5, -1.0, 0.3]) + 0.1 * rng.normal(size=150)
    ls = np.array([0.9, 1.3, 0.7])
    for form in ("expanded", "diff"):
        gp = O.obs_gp_fit(X, y, 1.4, ls, form=form)
        sk = GaussianProcessRegressor(ConstantKernel(1.4, "fixed") * RBF(ls, "fixed"), alpha=1e-2 + 1e-8, optimizer=None).fit(X, y)
        Xn = rng.normal(size=(40, 3))
        mu, sd = sk.predict(Xn, return_std=True)
        mu2, var2 = O.obs_gp_predict(gp, Xn)
        np.testing.assert_allclose(mu2, mu, atol=1e-10)
        np.testing.assert_allclose(var2, sd ** 2 + 1e-2, atol=1e-10)


def test_expanded_and_difference_distances_agree_on_moderate_coordinates():
    rng = np.random.default_rng(1)
    A, B = rng.normal(size=(20, 4)), rng.normal(size=(30, 4))
    for ls in (np.array([1.3]), np.array([0.5, 1.0, 2.0, 4.0])):
        np.testing.assert_allclose(O.scaled_sqdist(A, B, ls, "expanded"), O.scaled_sqdist(A, B, ls, "diff"), atol=1e-12)
    # large-magnitude coordinates (coral T ~ 2400): the expanded form loses ~1e-9 absolute
    A2, B2 = A + 2400.0, B + 2400.0
    err = np.abs(O.scaled_sqdist(A2, B2, np.array([1.0]), "expanded") - O.scaled_sqdist(A, B, np.array([1.0]), "diff")).max()
    assert 0 < err < 1e-6


def test_jitchol_retry_rule():
    A = np.array([[1.0, 1.0], [1.0, 1.0 - 1e-12]])           # not positive definite
    L, tries = O.jitchol(A)
    assert tries == 1                                         # mean(diag) * 1e-6 is enough
    np.testing.assert_allclose(L @ L.T, A + np.eye(2) * 1e-6 * np.mean(np.diag(A)), atol=1e-15)
    L0, t0 = O.jitchol(np.eye(3) * 2.0)
    assert t0 == 0
    with pytest.raises(np.linalg.LinAlgError):
        O.jitchol(np.array([[1.0, 2.0], [2.0, -1.0]]))


def test_expected_improvement_closed_form():
    mu = np.array([0.0, 0.5, -1.0, 3.0]); var = np.array([1.0, 0.25, 4.0, 1e-8]); best = 0.2
    sd = np.sqrt(var); u = (best - mu) / sd
    ref = sd * (u * scipy.stats.norm.cdf(u) + scipy.stats.norm.pdf(u))
    np.testing.assert_allclose(O.expected_improvement(mu, var, best, "min"), ref, rtol=1e-14)
    np.testing.assert_allclose(O.expected_improvement(mu, var, best, "max"), -ref, rtol=1e-14)   # reference :41
    assert np.all(ref >= 0)
    assert np.isnan(O.expected_improvement(np.array([0.0]), np.array([-1e-9]), 0.0)[0])           # no variance clip


def test_argmax_semantics():
    assert O.first_argmax([1.0, 3.0, 3.0, 2.0])[:2] == (1, 3.0)             # first maximum
    i, v, n = O.first_argmax([np.nan, -1.0, np.nan])
    assert (i, v, n) == (1, -1.0, 2)                                         # NaN = -inf, counted
    assert O.first_argmax([np.nan, np.nan])[0] == 0
    assert O.select_set([0.1, 0.7, 0.7])[0] == 1                             # CBO.py:275-276
    g = O.tensor_grid([np.array([0.0, 1.0]), np.array([10.0, 20.0, 30.0])])
    np.testing.assert_array_equal(g[4], [1.0, 20.0])                         # C order, last dimension fastest


def test_cost_forms():
    X = np.array([[1.0, -2.0], [0.5, 0.5]])
    np.testing.assert_array_equal(O.point_cost(X, [1, 10], False), [11.0, 11.0])
    np.testing.assert_array_equal(O.point_cost(X, [1, 10], True), [14.0, 12.0])


def test_prior_approaches_true_causal_effect_on_toy_graph():
    """do(Z) on X -> Z -> Y has no back-door path, so the do-prior of set ['Z'] is a GP regression of Y on Z and must
    track the shipped true effect cos z - exp(-z/20) inside the data range (loose, statistical)."""
    z = np.load(os.path.join(GOLD, "data", "toy_graph.npz"))
    cols = [str(c) for c in z["columns"]]
    obs = z["observations"][:400]
    Z, Y = obs[:, cols.index("Z")][:, None], obs[:, cols.index("Y")]
    gp = O.obs_gp_fit(Z, Y, 4.0, np.array([2.0]), form="diff")
    zq = np.linspace(0.0, 3.0, 16)[:, None]
    m, v = O.do_prior_factorised(gp, O.prior_factors(gp, Z, [0]), [0], zq)
    truth = np.cos(zq[:, 0]) - np.exp(-zq[:, 0] / 20.0)
    assert np.max(np.abs(m - truth)) < 0.6
    j = [str(n) for n in z["set1_names"]]
    assert j == ["Z"]
    np.testing.assert_allclose(z["set1_y"][:, 0], np.cos(z["set1_x"][:, 0]) - np.exp(-z["set1_x"][:, 0] / 20.0), atol=1e-10)


@pytest.mark.parametrize("config", ["toy", "complete"])
def test_oracle_reproduces_golden_vectors(config):
    """Guards the oracle against drift: recompute from the stored inputs and compare with the stored outputs."""
    z = np.load(os.path.join(GOLD, f"golden_{config}.npz"), allow_pickle=False)
    best = float(z["best"])
    for s in range(int(z["num_sets"])):
        k = f"set{s}_"
        X = np.hstack([z[k + "x_obs_int"], z[k + "x_obs_cond"]])
        d = z[k + "x_obs_int"].shape[1]
        gp = dict(X=X, variance=float(z[k + "s2"]), lengthscale=np.concatenate([z[k + "ls_int"], z[k + "ls_cond"]]), noise=1e-2,
                  alpha=z[k + "alpha_obs"], Kyinv=z[k + "kyinv"], form="diff")
        grid = [np.linspace(lo, hi, int(p)) for lo, hi, p in z[k + "grid_lo_hi_p"]]
        ref = O.sweep_set(gp, X, list(range(d)), z[k + "x_int"], z[k + "y_int"], grid, best, "min",
                          fix_costs=np.array([float(z[k + "cost_fix"])]), form="diff", precise_int=True)
        assert ref["idx"] == int(z[k + "idx"]) and ref["tries"] == int(z[k + "tries"])
        keep = z[k + "keep"]
        np.testing.assert_array_equal(ref["mI"], z[k + "mI"])       # interventional rows: extended precision both times
        np.testing.assert_array_equal(ref["vI"], z[k + "vI"])
        # kept candidates: stored from the extended-precision prior, recomputed here in float64 (its rounding noise is
        # the quadratic form's condition number times eps, see do_prior_factorised)
        scale = np.abs(z[k + "acq"]).max()
        np.testing.assert_allclose(ref["acq"][keep], z[k + "acq"], rtol=1e-7, atol=1e-9 * scale)
        np.testing.assert_allclose(ref["vg"][keep], z[k + "vg"], rtol=1e-8)


def test_sweep_set_direct_equals_factorised_small():
    kw, ora = make_case(9, N=40, d=2, c=1, n=6, p=(5, 4))
    best = float(np.min(kw["y_int"]))
    a = oracle_sweep(ora, best, prior="direct")
    b = oracle_sweep(ora, best, prior="factorised")
    assert a["idx"] == b["idx"]
    np.testing.assert_allclose(a["acq"], b["acq"], rtol=1e-8)
