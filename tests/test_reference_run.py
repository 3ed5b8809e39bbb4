"""Pins the oracle against the REFERENCE's own code.

tests/golden/reference_run_<config>.npz holds what the reference's modules (DoCalculus.compute_do, CausalRBF,
GaussianProcessFactory.create, CausalExpectedImprovement, Cost, find_current_global, CBO.select_next_intervention --
imported unmodified from /root/reference by tests/golden/make_reference_golden.py) returned on the inputs frozen in
tests/golden/golden_<config>.npz (BASELINE.json configs[0] toy_graph and configs[2] complete_graph on the full
100-points-per-dimension grid; configs[1] simplified_coral_graph with its ten 1e6-candidate sets sampled).  GPy / emukit / paramz themselves are not installable here and were replaced by the
stand-in under tests/golden/gpy_standin (its README lists what that leaves unpinned: GPy's own internals).

Two comparisons:
  * the oracle in its LITERAL form (one candidate at a time through the observational GP's predict, GPy's expanded
    distance) must reproduce the reference run to 1e-8 -- it is the same arithmetic in the same order;
  * the oracle in the form every other test uses (factorised prior, distances from coordinate differences) must agree
    under the 1e-6 parity rule, with identical integers (per-set argmax, jitter retries, NaN count, selected set).
"""
import os

import numpy as np
import pytest

from helpers import RTOL, sweep_errors
from oracle import cbo_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _set_inputs(z, s):
    k = f"set{s}_"
    X = np.hstack([z[k + "x_obs_int"], z[k + "x_obs_cond"]])
    d = z[k + "x_obs_int"].shape[1]
    ls = np.concatenate([z[k + "ls_int"], z[k + "ls_cond"]])
    grid = [np.linspace(lo, hi, int(p)) for lo, hi, p in z[k + "grid_lo_hi_p"]]
    return k, X, d, ls, grid


@pytest.mark.parametrize("config", ["toy", "complete"])
def test_default_oracle_form_agrees_with_reference_run(config):
    z = np.load(os.path.join(GOLD, f"golden_{config}.npz"), allow_pickle=False)
    r = np.load(os.path.join(GOLD, f"reference_run_{config}.npz"), allow_pickle=False)
    assert float(r["best"]) == float(z["best"])                    # find_current_global (utils.py:8-26)
    assert int(r["selected_set"]) == int(z["selected_set"])        # CBO.select_next_intervention (CBO.py:269-277)
    for s in range(int(z["num_sets"])):
        k = f"set{s}_"
        np.testing.assert_array_equal(r[k + "keep"], z[k + "keep"])
        assert int(r[k + "idx"]) == int(z[k + "idx"]), (config, s)
        assert int(r[k + "tries"]) == int(z[k + "tries"]) and int(r[k + "n_nan"]) == int(z[k + "n_nan"])
        assert float(r[k + "top2_gap"]) > 1e-6                     # no ties: the argmax comparison above is meaningful
        errs = sweep_errors({n: z[k + n] for n in ("mI", "vI", "mg", "vg", "mu", "var", "ei", "acq")}, r, k)
        for name, e in errs.items():
            assert e <= RTOL, f"{config} set {s} {name}: {e:.3e}"
        np.testing.assert_allclose(float(z[k + "val"]), float(r[k + "val"]), rtol=RTOL)


@pytest.mark.parametrize("config,sets", [("toy", None), ("complete", (0, 1, 2))])
def test_literal_oracle_form_reproduces_reference_run(config, sets):
    """Direct loop (DoCalculus.py:50-89) + GPy's expanded distances: same arithmetic as the reference run."""
    z = np.load(os.path.join(GOLD, f"golden_{config}.npz"), allow_pickle=False)
    r = np.load(os.path.join(GOLD, f"reference_run_{config}.npz"), allow_pickle=False)
    best = float(r["best"])
    for s in (range(int(z["num_sets"])) if sets is None else sets):
        k, X, d, ls, grid = _set_inputs(z, s)
        gp = O.obs_gp_fit(X, z[k + "y_obs"], float(z[k + "s2"]), ls if X.shape[1] > 1 else ls[:1], form="expanded")
        out = O.sweep_set(gp, X, list(range(d)), z[k + "x_int"], z[k + "y_int"], grid, best, "min",
                          fix_costs=np.array([float(z[k + "cost_fix"])]), prior="direct", form="expanded")
        keep = r[k + "keep"]
        assert out["idx"] == int(r[k + "idx"]) and out["tries"] == int(r[k + "tries"]) and out["n_nan"] == int(r[k + "n_nan"])
        np.testing.assert_allclose(out["mI"], r[k + "mI"], rtol=1e-9)
        np.testing.assert_allclose(out["vI"], r[k + "vI"], rtol=1e-9)
        np.testing.assert_allclose(out["mg"][keep], r[k + "mg"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out["vg"][keep], r[k + "vg"], rtol=1e-9)
        np.testing.assert_allclose(out["L"], r[k + "L"], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(out["mu"][keep], r[k + "mu"], rtol=1e-8, atol=1e-8 * np.abs(r[k + "mu"]).max())
        np.testing.assert_allclose(out["var"][keep], r[k + "var"], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(out["acq"][keep], r[k + "acq"], rtol=1e-8, atol=1e-8 * np.abs(r[k + "acq"]).max())
        np.testing.assert_allclose(out["val"], float(r[k + "val"]), rtol=1e-8)


def test_oracle_agrees_with_reference_run_on_the_coral_family():
    """simplified_coral_graph (BASELINE.json configs[1]: 25 exploration sets, d up to 3, coordinates ~2400; CoralGraph.py:54-70,
    163-184): the reference's own modules on the frozen inputs -- every candidate of the 1-D and 2-D sets, a seeded 4096-candidate
    sample plus the argmax neighbourhood of the 3-D sets (make_reference_golden.py) -- against the oracle at the same candidates.

    Two forms of the oracle's per-set GP: GPy's expanded distances (what the reference executes) must reproduce the run under
    the 1e-6 rule on EVERY set (measured: 3e-8); coordinate differences (what the CUDA path and every other test use) must do
    so wherever the reference's own rounding allows it -- on set TC (T ~ 2400, lengthscale 1) the expanded form loses 1e-9 in
    r^2, which the set's ill-conditioned fit turns into 7e-6 in the variance: there the difference form is held
    to three times the distance between the two forms, i.e. to the reference's own noise."""
    from helpers import form_distance, oracle_at_reference_points
    z = np.load(os.path.join(GOLD, "golden_simplified_coral.npz"), allow_pickle=False)
    r = np.load(os.path.join(GOLD, "reference_run_simplified_coral.npz"), allow_pickle=False)
    best = float(r["best"])
    assert best == float(z["best"]) and int(r["selected_set"]) == int(z["selected_set"])
    flat, sampled, noisy = [], 0, []
    val_scale = max(abs(float(r[f"set{s}_val"])) for s in range(int(z["num_sets"])))
    for s in range(int(z["num_sets"])):
        k = f"set{s}_"
        got_e, err_e = oracle_at_reference_points(z, r, s, best, "expanded", val_scale)
        got_d, err_d = oracle_at_reference_points(z, r, s, best, "diff", val_scale)
        for name, e in err_e.items():
            assert e <= RTOL, f"expanded form, set {s} ({z[k + 'name']}) {name}: {e:.3e}"
        for name, e in err_d.items():
            if e > RTOL:
                noisy.append((str(z[k + "name"]), name, float(e)))
                assert e <= 3.0 * (err_e[name] + form_distance(got_d, got_e, k, val_scale)[name]), (s, name, e)
        assert int(r[k + "tries"]) == int(z[k + "tries"]) == got_d["tries"] == got_e["tries"]
        sampled += int(r[k + "sampled"])
        # (a set whose whole acquisition sits in the far tail of EI -- 1e-66 against 30 for the winning set -- is compared on
        # the scale of the trial's acquisition values, like `ei` / `acq` inside sweep_errors)
        np.testing.assert_allclose(float(z[k + "val"]), float(r[k + "val"]), rtol=RTOL, atol=RTOL * 1e-6 * val_scale)
        if min(float(r[k + "top2_gap"]), float(z[k + "top2_gap"])) > 1e-9:
            assert int(r[k + "idx"]) == int(z[k + "idx"]), (s, str(z[k + "name"]))      # bit-exact argmax wherever it is unique
        else:
            flat.append(str(z[k + "name"]))
            if int(r[k + "idx"]) != int(z[k + "idx"]):   # a flat maximum: both picks must carry the same acquisition value
                keep = r[k + "keep"]
                j = int(np.nonzero(keep == int(r[k + "idx"]))[0][0])
                assert abs(got_d["acq"][j] - float(z[k + "val"])) <= 1e-9 * abs(float(z[k + "val"]))
    assert sampled == 10 and len(flat) <= 12, (sampled, flat)
    assert {n for n, _, _ in noisy} <= {"TC"}, noisy          # the only set where the reference's distance rounding shows


def test_standin_is_not_imported_by_product_or_tests_at_run_time():
    """The stand-in and the reference are build-container tools: nothing else may depend on them."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for base in ("cbo_with_oop_b200", "src", "oracle", "bench.py", "__graft_entry__.py", "tests"):
        path = os.path.join(root, base)
        files = [path] if os.path.isfile(path) else [os.path.join(dp, f) for dp, _, fs in os.walk(path) for f in fs if f.endswith(".py")]
        for f in files:
            if os.sep + "golden" + os.sep in f or f.endswith("test_reference_run.py"):
                continue
            for line in open(f):
                t = line.strip()
                if t.startswith(("import GPy", "from GPy", "import emukit", "from emukit", "import paramz", "from paramz")) \
                        or ("sys.path" in t and ("gpy_standin" in t or "reference" in t)):
                    offenders.append(os.path.relpath(f, root))
    assert not offenders, offenders
