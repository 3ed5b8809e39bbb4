"""GPU parity: every stage of the CUDA path (through the C ABI, via SweepEngine) against the CPU oracle on
identical seeded inputs.  Integer results (selected set, selected index, jitter retries, NaN count) must be
bit-exact; floating point must agree within 1e-6 relative (north_star), with the floor rule of helpers.rel_err
where the reference value is a cancelled difference of O(1) terms."""
import numpy as np
import pytest

from helpers import BACKWARD_TOL, MODERATE_COND, RTOL, fit_errors, floor_count, make_case, oracle_sweep, rel_err
from oracle import cbo_oracle as O

pytestmark = pytest.mark.gpu


def _engine(kws, **kw):
    from cbo_with_oop_b200.engine import SetProblem, SweepEngine
    return SweepEngine([SetProblem(**k) for k in kws], keep=("mu", "var", "ei", "acq"), **kw)


def _check_set(eng, g, ref, ora, report):
    gp = ora["gp"]
    if ora["causal"]:
        f = O.prior_factors(gp, ora["cond"], ora["cols"])
        scale_M = np.abs(f["M"]).max()
        report["M"] = rel_err(eng.fetch("M", g), f["M"], 1e-6 * scale_M).max()
        report["w"] = rel_err(eng.fetch("w", g), f["w"], 1e-6 * np.abs(f["w"]).max()).max()
        report["m_int"] = rel_err(eng.fetch("m_int", g), ref["mI"], 1e-6).max()
        report["v_int"] = rel_err(eng.fetch("v_int", g), ref["vI"], 1e-6).max()
        report["m"] = rel_err(eng.fetch("m", g), ref["mg"], 1e-6).max()
        report["v"] = rel_err(eng.fetch("v", g), ref["vg"], 1e-6).max()
        kdiag = 1.0 + ref["vg"]
    else:
        kdiag = np.ones_like(ref["mu"])
    # the fit's own inputs (the kernel's prior at x_int), so that the residuals measure the factorisation and the solves only
    own = (eng.fetch("m_int", g), eng.fetch("v_int", g)) if ora["causal"] else (None, None)
    fe = fit_errors(eng.fetch("L", g), eng.fetch("alpha", g), ora["XI"], ora["yI"], own[0], own[1], ref["L"], ref["alpha"])
    assert fe["L_backward"] <= BACKWARD_TOL and fe["alpha_backward"] <= BACKWARD_TOL, fe
    if fe["cond"] < MODERATE_COND:
        report["L"], report["alpha"] = fe["L_forward"], fe["alpha_forward"]
    report["fit_cond_pts"] = fe["cond"]
    report["mu"] = rel_err(eng.fetch("mu", g), ref["mu"], 1e-4).max()
    report["var"] = rel_err(eng.fetch("var", g), ref["var"], 1e-4 * kdiag).max()
    report["var_floor_pts"] = floor_count(ref["var"] / kdiag, 1e-4)
    ei_scale = max(np.nanmax(np.abs(ref["ei"])), 1e-300)
    report["ei"] = np.nanmax(rel_err(eng.fetch("ei", g), ref["ei"], 1e-6 * ei_scale))
    report["acq"] = np.nanmax(rel_err(eng.fetch("acq", g), ref["acq"], 1e-6 * ei_scale))
    return report


CASES = [
    dict(seed=1, N=96, d=1, c=0, n=7, p=(33,)),
    dict(seed=2, N=130, d=1, c=2, n=10, p=(100,)),
    dict(seed=3, N=200, d=2, c=2, n=12, p=(17, 23)),
    dict(seed=4, N=257, d=3, c=3, n=16, p=(9, 10, 11), S_mc=77),
    dict(seed=5, N=64, d=2, c=1, n=5, p=(16, 8), ard=False, cost_variable=True),
    dict(seed=6, N=300, d=3, c=5, n=33, p=(12, 12, 12), S_mc=150),
    dict(seed=7, N=128, d=2, c=0, n=9, p=(31, 5)),
    dict(seed=8, N=100, d=2, c=2, n=10, p=(21, 21), duplicate_train_point=True),
    # 16 < n <= 48: |L^-1 k*|^2 through the DMMA product with L^-1 (sweep_mma_kernel), separable tables and general form,
    # partial tiles, n on and off the 8-row block / k4-step boundaries, a candidate on an interventional row
    dict(seed=61, N=96, d=2, c=1, n=17, p=(40, 45)),
    dict(seed=62, N=120, d=3, c=2, n=24, p=(20, 20, 37)),
    dict(seed=63, N=96, d=2, c=2, n=32, p=(50, 101), cost_variable=True),
    dict(seed=64, N=130, d=3, c=1, n=45, p=(10, 12, 100), duplicate_train_point=True),
    # (1-D sets this size need room: 48 points inside [-2, 2] give a fit of condition 3e9, where two backward-stable CPU
    # solvers already differ by 2e-5 in mu)
    dict(seed=71, N=64, d=1, c=2, n=48, p=(2000,), lo=-60.0, hi=60.0),
    dict(seed=66, N=64, d=1, c=1, n=41, p=(200,), lo=-30.0, hi=30.0),
    dict(seed=67, N=80, d=2, c=1, n=30, p=(300, 7), duplicate_train_point=True),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"seed{c['seed']}")
@pytest.mark.parametrize("task", ["min", "max"])
def test_single_set_all_stages(cuda_engine_ready, case, task):
    kw, ora = make_case(**case)
    best = float(np.min(kw["y_int"]) if task == "min" else np.max(kw["y_int"]))
    ref = oracle_sweep(ora, best, task)
    eng = _engine([kw])
    out = eng.sweep(best, task)
    rep = _check_set(eng, 0, ref, ora, {})
    print(case["seed"], task, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in rep.items()})
    info = eng.fetch("fit_info", 0)
    assert info[1] == 0 and info[0] == ref["tries"]
    if case["n"] <= 48:     # L^-T in the strict upper triangle of the factor (input of K3's tensor-pipe path)
        L, W = eng.fetch("L", 0), eng.fetch("Linv", 0)
        assert np.abs(W @ L - np.eye(case["n"])).max() <= 1e-9 * np.linalg.cond(L)
    for k, v in rep.items():
        if k.endswith("_pts"):
            continue
        assert v <= RTOL, f"{k}: {v:.3e} > {RTOL}"
    # bit-exact selection
    srt = np.sort(ref["acq"][~np.isnan(ref["acq"])])
    gap = (srt[-1] - srt[-2]) / abs(srt[-1]) if srt.size > 1 and srt[-1] != 0 else np.inf
    assert out.set == 0
    assert out.index == ref["idx"], f"argmax {out.index} != oracle {ref['idx']} (top-2 relative gap {gap:.2e})"
    assert out.n_nan == ref["n_nan"]
    np.testing.assert_allclose(out.value, ref["val"], rtol=RTOL)
    np.testing.assert_array_equal(out.x, ref["x"])


def test_non_causal_set(cuda_engine_ready):
    kw, ora = make_case(seed=11, N=64, d=2, c=1, n=10, p=(25, 20), causal=False)
    best = float(np.min(kw["y_int"]))
    ref = oracle_sweep(ora, best, "min")
    eng = _engine([kw])
    out = eng.sweep(best, "min")
    rep = _check_set(eng, 0, ref, ora, {})
    for k, v in rep.items():
        if not k.endswith("_pts"):
            assert v <= RTOL, f"{k}: {v:.3e}"
    assert out.index == ref["idx"]


def test_direct_form_agrees(cuda_engine_ready):
    """The reference-faithful loop form (DoCalculus.py:50-89 restated) against the CUDA factorised path."""
    kw, ora = make_case(seed=21, N=80, d=2, c=2, n=6, p=(6, 7))
    best = float(np.min(kw["y_int"]))
    ref = oracle_sweep(ora, best, "min", prior="direct")
    eng = _engine([kw])
    out = eng.sweep(best, "min")
    assert rel_err(eng.fetch("m", 0), ref["mg"], 1e-6).max() <= RTOL
    assert rel_err(eng.fetch("v", 0), ref["vg"], 1e-6).max() <= RTOL
    assert out.index == ref["idx"]


def test_multi_set_selection(cuda_engine_ready):
    """All exploration sets in one batched pass; first set attaining the maximum wins (CBO.py:275-276)."""
    specs = [dict(seed=31, N=96, d=1, c=1, n=8, p=(50,)), dict(seed=32, N=140, d=2, c=2, n=10, p=(20, 20)),
             dict(seed=33, N=200, d=3, c=1, n=12, p=(10, 10, 10)), dict(seed=34, N=70, d=2, c=0, n=9, p=(30, 9)),
             dict(seed=35, N=90, d=1, c=2, n=10, p=(64,), causal=False)]
    cases = [make_case(**s) for s in specs]
    best = float(min(np.min(k["y_int"]) for k, _ in cases))
    refs = [oracle_sweep(o, best, "min") for _, o in cases]
    eng = _engine([k for k, _ in cases])
    out = eng.sweep(best, "min")
    vals = np.array([r["val"] for r in refs])
    np.testing.assert_allclose(out.set_values, vals, rtol=RTOL)
    np.testing.assert_array_equal(out.set_indices, [r["idx"] for r in refs])
    s_ref, _ = O.select_set(vals)
    assert out.set == s_ref and out.index == refs[s_ref]["idx"]
    # duplicate the winning set in front of itself: the FIRST copy must be selected
    order = [s_ref] + list(range(len(cases)))
    eng2 = _engine([cases[i][0] for i in order])
    out2 = eng2.sweep(best, "min")
    assert out2.set == 0 and out2.index == refs[s_ref]["idx"]
    assert out2.set_values[0] == out2.set_values[1 + s_ref]


def test_refresh_after_intervention(cuda_engine_ready):
    """Post-intervention trial: cached prior, one set refit with an appended row (Monitor.py:148-160,
    CBO.py:224-235) equals a from-scratch sweep of the updated problem."""
    cases = [make_case(seed=41, N=96, d=2, c=1, n=8, p=(15, 15)), make_case(seed=42, N=110, d=1, c=2, n=9, p=(40,))]
    best = float(min(np.min(k["y_int"]) for k, _ in cases))
    eng = _engine([k for k, _ in cases])
    out = eng.sweep(best, "min")
    g = out.set
    kw, ora = cases[g]
    x_new = np.vstack([kw["x_int"], out.x[None, :]])
    y_new = np.append(kw["y_int"], best - 0.3)
    eng.set_interventional(g, x_new, y_new)
    out_r = eng.refresh(best - 0.3, "min", refit=[g])
    ora2 = dict(ora, XI=x_new, yI=y_new)
    refs = [oracle_sweep(ora2 if i == g else cases[i][1], best - 0.3, "min") for i in range(2)]
    np.testing.assert_allclose(out_r.set_values, [r["val"] for r in refs], rtol=RTOL)
    np.testing.assert_array_equal(out_r.set_indices, [r["idx"] for r in refs])
    # the appended row alone was evaluated (cbo_set_desc.int_row_begin); the earlier rows kept their prior bit for bit
    m_first = eng.fetch("m_int", g)
    np.testing.assert_allclose(m_first, refs[g]["mI"], rtol=RTOL)
    # a second appended row, then a rewritten row (no longer an append: every row is re-evaluated)
    lo = np.array([t[0] for t in kw["grid"]]); hi = np.array([t[-1] for t in kw["grid"]])
    x3 = np.vstack([x_new, (0.25 * lo + 0.75 * hi)[None, :]])
    y3 = np.append(y_new, best + 0.1)
    eng.set_interventional(g, x3, y3)
    out3 = eng.refresh(best - 0.3, "min", refit=[g])
    np.testing.assert_array_equal(eng.fetch("m_int", g)[:len(m_first)], m_first)
    x4 = x3.copy(); x4[1] = 0.5 * (lo + hi) + 0.01
    eng.set_interventional(g, x4, y3)
    out4 = eng.refresh(best - 0.3, "min", refit=[g])
    for xx, oo in ((x3, out3), (x4, out4)):
        refs = [oracle_sweep(dict(ora, XI=xx, yI=y3) if i == g else cases[i][1], best - 0.3, "min") for i in range(2)]
        np.testing.assert_allclose(oo.set_values, [r["val"] for r in refs], rtol=RTOL)
        np.testing.assert_array_equal(oo.set_indices, [r["idx"] for r in refs])
    np.testing.assert_allclose(eng.fetch("m_int", g), refs[g]["mI"], rtol=RTOL)
    np.testing.assert_allclose(eng.fetch("v_int", g), refs[g]["vI"], rtol=RTOL)


def test_jitter_retry_and_nan_policy(cuda_engine_ready):
    """Two identical interventional rows plus a third almost on top: the Gram is singular up to the 1e-8
    jitter; whatever the oracle's jitchol needs, the kernel must need too, and NaN acquisitions are counted."""
    kw, ora = make_case(seed=51, N=64, d=1, c=1, n=6, p=(41,), causal=False)
    kw["x_int"][1] = kw["x_int"][0]
    kw["y_int"][1] = kw["y_int"][0]
    ora["XI"], ora["yI"] = kw["x_int"], kw["y_int"]
    best = float(np.min(kw["y_int"]))
    ref = oracle_sweep(ora, best, "min")
    eng = _engine([kw])
    out = eng.sweep(best, "min")
    info = eng.fetch("fit_info", 0)
    assert info[0] == ref["tries"] and info[1] == 0
    assert out.n_nan == ref["n_nan"]
    assert out.index == ref["idx"]


def test_combine_kernel_matches_rule(cuda_engine_ready):
    """K4 on emulated ranks: gathered[rank][set] -> per-set bests and the global pick."""
    import ctypes as C
    import torch
    from cbo_with_oop_b200 import _lib
    from cbo_with_oop_b200._lib import SetBest, SweepResult
    lib = _lib.load()
    R, S = 4, 6
    rng = np.random.default_rng(0)
    g = (SetBest * (R * S))()
    vals = rng.standard_normal((R, S))
    vals[1, 2] = vals[3, 2] = 5.0   # tie inside a set: lowest index wins
    vals[:, 4] = -np.inf            # a set nobody evaluated
    vals[0, 0] = vals[1, 2]         # tie across sets: lowest set wins
    vals[:, 0] = np.minimum(vals[:, 0], 5.0)
    for r in range(R):
        for s in range(S):
            e = g[r * S + s]
            e.value, e.index, e.n_nan = vals[r, s], (-1 if s == 4 else 1000 * (R - r) + s), r
    dev = torch.device("cuda:0")
    dg = torch.frombuffer(bytearray(bytes(g)), dtype=torch.uint8).to(dev)
    dsb = torch.empty(S * C.sizeof(SetBest), dtype=torch.uint8, device=dev)
    dres = torch.empty(C.sizeof(SweepResult), dtype=torch.uint8, device=dev)
    _lib.check(lib.cbo_argmax_combine(dg.data_ptr(), R, S, dsb.data_ptr(), dres.data_ptr(), None), "combine")
    torch.cuda.synchronize()
    sb = (SetBest * S).from_buffer_copy(dsb.cpu().numpy().tobytes())
    res = SweepResult.from_buffer_copy(dres.cpu().numpy().tobytes())
    for s in range(S):
        col = vals[:, s]
        if s == 4:
            assert sb[s].index == -1
            continue
        cand = [(-col[r], 1000 * (R - r) + s) for r in range(R)]
        v, i = min(cand)
        assert sb[s].value == -v and sb[s].index == i
        assert sb[s].n_nan == sum(range(R))
    assert res.set == 0 and res.value == 5.0
    assert res.n_nan == (S) * sum(range(R))


@pytest.mark.parametrize("S_mc", [40, 333])
def test_prior_precompute_of_a_set_that_fills_the_gpu(cuda_engine_ready, S_mc):
    """N = 2200 (18 row blocks, 171 lower-triangle tiles >= the 148 SMs): K1a's SYRK takes the persistent tensor-map TMA
    pipeline (dmma_tma_tile.cuh) instead of the one-tile-per-CTA kernel.  S_mc = 40 is a product of 3 slabs -- shorter
    than the 6-stage ring -- and 333 wraps it several times.  M, w and the prior at the interventional rows against the oracle."""
    kw, ora = make_case(seed=41, N=2200, d=1, c=2, n=6, p=(50,), S_mc=S_mc)
    eng = _engine([kw])
    best = float(np.min(kw["y_int"]))
    out = eng.sweep(best, "min")
    ref = oracle_sweep(ora, best, "min")
    f = O.prior_factors(ora["gp"], ora["cond"], ora["cols"])
    assert rel_err(eng.fetch("M", 0), f["M"], 1e-6 * np.abs(f["M"]).max()).max() <= RTOL
    assert rel_err(eng.fetch("w", 0), f["w"], 1e-6 * np.abs(f["w"]).max()).max() <= RTOL
    assert rel_err(eng.fetch("m", 0), ref["mg"], 1e-6).max() <= RTOL
    assert rel_err(eng.fetch("v", 0), ref["vg"], 1e-6).max() <= RTOL
    assert out.index == ref["idx"]
