"""Full-size check on BASELINE.json configs[4] (one exploration set of the synthetic scaled sweep: 1e6-point grid,
1e4 observational samples, 32 interventional rows).  The oracle cannot sweep 1e6 candidates at N = 1e4 in test time,
so parity is taken on a seeded sample of candidates, and the whole grid is held to size-independent properties:
  * noise <= v(x) <= s2 + noise        (v is a mean of GP predictive variances that include the noise)
  * grid path == explicit-point path   (the interventional rows lie on the grid: m, v there must equal m_int, v_int)
  * interpolation                      (noise 1e-10: mu == y_int and var ~ 1e-8 at the interventional rows)
  * argmax                             (reported value is the maximum of the acquisition array, index is its first maximiser)
  * partition invariance               (the two slices of a 2-way split reproduce the single-GPU arrays bit for bit)
"""
import numpy as np
import pytest

from helpers import RTOL, rel_err
from oracle import cbo_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("set_index,n_sample", [(3, 192), (11, 512)])
def test_config5_one_set_full_size(cuda_engine_ready, set_index, n_sample):
    import torch
    from cbo_with_oop_b200.engine import SweepEngine
    from cbo_with_oop_b200.synthetic import scaled_set
    pr = scaled_set(set_index, n_obs=10_000, p=100, d=3, c=3, n_int=32, device="cuda:0")
    best = float(pr.y_int.min())
    eng = SweepEngine([pr], keep=("mu", "var", "ei", "acq"))
    out = eng.sweep(best, "min")
    m, v, mu, var, acq = (eng.fetch(k, 0) for k in ("m", "v", "mu", "var", "acq"))
    G = pr.g_total
    assert m.shape == (G,) and G == 10 ** 6

    # ---- properties over the whole grid
    assert np.all(np.isfinite(m)) and np.all(np.isfinite(v))
    assert v.min() >= pr.noise * (1 - 1e-9) and v.max() <= pr.s2 + pr.noise + 1e-12
    shape = [len(t) for t in pr.grid]
    flat = np.ravel_multi_index([np.searchsorted(pr.grid[k], pr.x_int[:, k]) for k in range(3)], shape)
    np.testing.assert_array_equal(np.stack([pr.grid[k][np.unravel_index(flat, shape)[k]] for k in range(3)], 1), pr.x_int)
    assert rel_err(m[flat], eng.fetch("m_int", 0), 1e-6).max() <= 1e-10      # same kernel, different tiling / split
    assert rel_err(v[flat], eng.fetch("v_int", 0), 1e-6).max() <= 1e-10
    assert np.abs(mu[flat] - pr.y_int).max() <= 1e-6 * max(1.0, np.abs(pr.y_int).max())
    assert np.all(var[flat] > 0) and var[flat].max() < 1e-6
    assert out.n_nan == int(np.isnan(acq).sum())
    a = np.where(np.isnan(acq), -np.inf, acq)
    assert out.index == int(np.argmax(a)) and out.value == a[out.index]

    # ---- oracle parity on a seeded sample of candidates (plus the selected one)
    rng = np.random.default_rng(0)
    idx = np.unique(np.concatenate([rng.choice(G, n_sample, replace=False), flat[:8], [out.index]]))
    Xs = np.stack([pr.grid[k][np.unravel_index(idx, shape)[k]] for k in range(3)], 1)
    X = np.hstack([pr.x_obs_int, pr.x_obs_cond])
    gp = dict(X=X, variance=pr.s2, lengthscale=np.concatenate([pr.ls_int, pr.ls_cond]), noise=pr.noise, alpha=pr.alpha_obs,
              Kyinv=pr.kyinv, form="diff")
    f = O.prior_factors(gp, X, [0, 1, 2])
    mI, vI = O.do_prior_factorised(gp, f, [0, 1, 2], pr.x_int)
    ms, vs = O.do_prior_factorised(gp, f, [0, 1, 2], Xs)
    post = O.posterior_fit(pr.x_int, pr.y_int, mI, vI, form="diff")
    mus, vars_ = O.posterior_predict(post, Xs, ms, vs)
    acqs = O.expected_improvement(mus, vars_, best, "min") / pr.cost_fix
    assert rel_err(eng.fetch("m_int", 0), mI, 1e-6).max() <= RTOL and rel_err(eng.fetch("v_int", 0), vI, 1e-6).max() <= RTOL
    assert rel_err(m[idx], ms, 1e-6).max() <= RTOL
    assert rel_err(v[idx], vs, 1e-6).max() <= RTOL
    assert rel_err(mu[idx], mus, max(1e-4, 1e-3 * np.abs(mus).max())).max() <= RTOL
    assert rel_err(var[idx], vars_, 1e-4 * (1 + vs)).max() <= RTOL
    assert rel_err(acq[idx], acqs, 1e-6 * np.abs(acqs).max()).max() <= RTOL
    assert np.nanmax(acqs) <= out.value * (1 + 1e-9)

    # ---- partition invariance: the two ranks' slices of a 2-way split, run one after the other on this GPU through
    # the stage calls (no process group needed before the all-gather), reproduce the unsplit arrays bit for bit
    from cbo_with_oop_b200.partition import SetSize, partition
    sl = partition([SetSize(G, 10_000, 32)], 2, snap=0.0)
    assert sl[0][0][1] + sl[1][0][1] == G and sl[0][0][1] % 128 == 0 and sl[1][0][0] == sl[0][0][1]
    del eng
    torch.cuda.empty_cache()
    for r in range(2):
        e = SweepEngine([pr], rank=r, world_size=2, keep=("acq",))
        gb, gc = e.slices[0]
        assert (gb, gc) == tuple(sl[r][0])
        e.build_tables(); e.prior_precompute(); e.prior_eval(1); e.posterior_fit(); e.prior_eval(0)
        e._sweep_local(best, "min")
        torch.cuda.synchronize()
        np.testing.assert_array_equal(e.fetch("m", 0), m[gb:gb + gc])
        np.testing.assert_array_equal(e.fetch("v", 0), v[gb:gb + gc])
        np.testing.assert_array_equal(e.fetch("acq", 0), acq[gb:gb + gc])
        del e
        torch.cuda.empty_cache()


def test_small_tensor_grid_at_full_n_obs(cuda_engine_ready):
    """A 2048-candidate tensor grid at N = 10^4 (16 full tiles: too few for M's triangle to be cut into segments, so the column
    blocks are dealt to several items per tile) with a ragged last column block (10^4 = 78 * 128 + 16): the item that does NOT
    own the last block must not wait for it (a round-1 deadlock: the kernel trapped).  Checked against the explicit-point path."""
    import torch
    from cbo_with_oop_b200.engine import SweepEngine
    from cbo_with_oop_b200.synthetic import scaled_set
    pr = scaled_set(3, n_obs=10_000, p=16, d=3, c=3, n_int=32, device_fit=True)
    pr.grid = [np.linspace(-2.0, 2.0, pk) for pk in (8, 16, 16)]
    eng = SweepEngine([pr])
    out = eng.sweep(float(np.min(pr.y_int)), "min")
    torch.cuda.synchronize()
    m, v = eng.fetch("m", 0), eng.fetch("v", 0)
    assert np.all(np.isfinite(m)) and np.all(np.isfinite(v)) and 0 <= out.index < 2048
    flat = np.arange(0, 2048, 7)
    ii = np.unravel_index(flat, (8, 16, 16))
    X = np.stack([pr.grid[k][ii[k]] for k in range(3)], axis=1)
    pts = eng.evaluate_points(0, X, stages="prior")
    assert rel_err(m[flat], pts["m"], 1e-6).max() <= 1e-9 and rel_err(v[flat], pts["v"], 1e-6).max() <= 1e-9
