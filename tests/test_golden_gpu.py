"""GPU parity against the committed golden fixtures.  Inputs: tests/golden/golden_*.npz (make_golden.py, from the
shipped data of the reference: BASELINE.json configs 1-3, and config 4 with synthetic observations).  Expected outputs:
  * "oracle":        the oracle's, stored in the same file;
  * "reference_run": what the REFERENCE's own modules returned on those inputs (tests/golden/reference_run_*.npz, written
                     by make_reference_golden.py in the build container; GPy itself replaced by tests/golden/gpy_standin).
All exploration sets of a config go through ONE batched sweep; integer results must be bit-exact."""
import json
import os

import numpy as np
import pytest

from helpers import BACKWARD_TOL, MODERATE_COND, RTOL, fit_errors, rel_err, sweep_errors

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_problems(z):
    from cbo_with_oop_b200.engine import SetProblem
    from cbo_with_oop_b200.obs_gp import fit_state
    problems = []
    for s in range(int(z["num_sets"])):
        k = f"set{s}_"
        xi, xc = z[k + "x_obs_int"], z[k + "x_obs_cond"]
        ls_int, ls_cond, s2 = z[k + "ls_int"], z[k + "ls_cond"], float(z[k + "s2"])
        if k + "kyinv" in z:
            alpha, kyinv = z[k + "alpha_obs"], z[k + "kyinv"]
        else:
            alpha, kyinv = fit_state(np.hstack([xi, xc]), z[k + "y_obs"], s2, np.concatenate([ls_int, ls_cond]), 1e-2)
            np.testing.assert_allclose(alpha, z[k + "alpha_obs"], rtol=1e-8, atol=1e-10 * np.abs(z[k + "alpha_obs"]).max())
        grid = [np.linspace(lo, hi, int(p)) for lo, hi, p in z[k + "grid_lo_hi_p"]]
        problems.append(SetProblem(x_obs_int=xi, x_obs_cond=xc, mc_cond=xc, alpha_obs=alpha, kyinv=kyinv, ls_int=ls_int,
                                   ls_cond=ls_cond, s2=s2, grid=grid, x_int=z[k + "x_int"], y_int=z[k + "y_int"],
                                   cost_fix=float(z[k + "cost_fix"]), name=str(z[k + "name"])))
    return problems


TIE_GAP = 1e-9


def oracle_acq_at(z, s, x, best):
    """Oracle acquisition of set s at ONE candidate x, from the inputs stored in the fixture."""
    from oracle import cbo_oracle as O
    k = f"set{s}_"
    X = np.hstack([z[k + "x_obs_int"], z[k + "x_obs_cond"]])
    d = z[k + "x_obs_int"].shape[1]
    from cbo_with_oop_b200.obs_gp import fit_state
    s2, ls = float(z[k + "s2"]), np.concatenate([z[k + "ls_int"], z[k + "ls_cond"]])
    kyinv = z[k + "kyinv"] if k + "kyinv" in z else fit_state(X, z[k + "y_obs"], s2, ls, 1e-2)[1]
    gp = dict(X=X, variance=s2, lengthscale=ls, noise=1e-2, alpha=z[k + "alpha_obs"], Kyinv=kyinv, form="diff")
    f = O.prior_factors(gp, X, list(range(d)))
    post = O.posterior_fit(z[k + "x_int"], z[k + "y_int"], z[k + "mI"], z[k + "vI"], form="diff")
    m, v = O.do_prior_factorised(gp, f, list(range(d)), x[None, :])
    mu, var = O.posterior_predict(post, x[None, :], m, v)
    return float(O.expected_improvement(mu, var, best, "min")[0] / float(z[k + "cost_fix"]))


@pytest.mark.parametrize("config,expected", [("toy", "oracle"), ("complete", "oracle"), ("simplified_coral", "oracle"),
                                             ("coral_synth", "oracle"), ("toy", "reference_run"),
                                             ("complete", "reference_run"), ("simplified_coral", "reference_run")])
def test_golden_config(cuda_engine_ready, config, expected):
    inputs = np.load(os.path.join(GOLD, f"golden_{config}.npz"), allow_pickle=False)
    z = inputs if expected == "oracle" else np.load(os.path.join(GOLD, f"reference_run_{config}.npz"), allow_pickle=False)
    assert float(z["best"]) == float(inputs["best"]) and int(z["num_sets"]) == int(inputs["num_sets"])
    from cbo_with_oop_b200.engine import SweepEngine
    problems = load_problems(inputs)
    eng = SweepEngine(problems, keep=("mu", "var", "ei", "acq"))
    out = eng.sweep(float(z["best"]), str(z["task"]))
    worst, tie_sets, floored = {}, [], {"var": 0, "mu": 0, "kept": 0}
    # per-set maxima deep in the tail of EI (1e-66 next to 30 for the winning set) are compared on the trial's scale
    val_scale = max(abs(float(z[f"set{s}_val"])) for s in range(len(problems)))
    for s in range(len(problems)):
        k = f"set{s}_"
        keep = z[k + "keep"]
        info = eng.fetch("fit_info", s)
        assert info[1] == 0 and info[0] == int(z[k + "tries"])
        XI, yI = inputs[k + "x_int"], inputs[k + "y_int"]
        fe = fit_errors(eng.fetch("L", s), eng.fetch("alpha", s), XI, yI, eng.fetch("m_int", s), eng.fetch("v_int", s),
                        z[k + "L"], z[k + "alpha"])
        assert fe["L_backward"] <= BACKWARD_TOL and fe["alpha_backward"] <= BACKWARD_TOL, (config, s, fe)
        worst["fit_cond"] = max(worst.get("fit_cond", 0.0), fe["cond"])
        got = {"mI": eng.fetch("m_int", s), "vI": eng.fetch("v_int", s), "mg": eng.fetch("m", s)[keep],
               "vg": eng.fetch("v", s)[keep], "mu": eng.fetch("mu", s)[keep], "var": eng.fetch("var", s)[keep],
               "ei": eng.fetch("ei", s)[keep], "acq": eng.fetch("acq", s)[keep]}
        errs = sweep_errors(got, z, k, ei_scale_floor=1e-6 * val_scale if expected == "reference_run" else 0.0)
        # how many kept candidates are judged against a floor instead of their own magnitude (helpers.sweep_errors)
        floored["kept"] += len(keep)
        floored["var"] += int(np.sum(np.abs(z[k + "var"]) < 1e-4 * (1.0 + z[k + "vg"])))
        floored["mu"] += int(np.sum(np.abs(z[k + "mu"]) < max(1e-4, 1e-3 * np.abs(z[k + "mu"]).max())))
        errs["L"] = fe["L_forward"] if fe["cond"] < MODERATE_COND else 0.0
        errs["alpha"] = fe["alpha_forward"] if fe["cond"] < MODERATE_COND else 0.0
        tol = {name: RTOL for name in errs}
        if expected == "reference_run" and config == "simplified_coral":
            # the reference's expanded distances carry their own rounding noise at coral's coordinates (~2400): where the
            # oracle's two distance forms differ by more than the tolerance (set TC), the CUDA path -- coordinate differences --
            # is held to three times that distance (tests/test_reference_run.py pins both forms against the run)
            from helpers import form_distance, oracle_at_reference_points
            o_d, _ = oracle_at_reference_points(inputs, z, s, float(z["best"]), "diff", val_scale)
            o_e, e_e = oracle_at_reference_points(inputs, z, s, float(z["best"]), "expanded", val_scale)
            dist = form_distance(o_d, o_e, k, val_scale)
            tol.update({name: max(RTOL, 3.0 * (dist[name] + float(e_e[name]))) for name in dist})
            noisy = [name for name in dist if tol[name] > RTOL]
            assert not noisy or str(z[k + "name"]) == "TC", (str(z[k + "name"]), noisy)
        for name, e in errs.items():
            worst[name] = max(worst.get(name, 0.0), float(e))
            assert e <= tol[name], f"{config} set {s} ({z[k + 'name']}) {name}: {e:.3e}; cond {fe['cond']:.2e}; all {errs}"
        # bit-exact selection -- unless the oracle's own top two candidates tie to rounding (flat acquisition far from
        # every interventional row): then any member of the tie set is a correct argmax and the GPU's pick must be one
        # (a sampled reference run sees the gap between its sample's two best candidates; the full-grid gap is the oracle's)
        gap = min(float(z[k + "top2_gap"]), float(inputs[k + "top2_gap"]))
        if out.set_indices[s] != int(z[k + "idx"]):
            assert gap <= TIE_GAP, f"{config} set {s}: argmax {out.set_indices[s]} != golden {int(z[k + 'idx'])} (top-2 gap {gap:.1e})"
            shape = [len(t) for t in problems[s].grid]
            ii = np.unravel_index(int(out.set_indices[s]), shape)
            x = np.array([problems[s].grid[a][ii[a]] for a in range(len(shape))])
            a_ref = oracle_acq_at(inputs, s, x, float(z["best"]))
            # (on the scale of the trial's acquisition values: a plateau at 1e-106 is compared like EI itself, see val_scale)
            assert abs(a_ref - float(z[k + "val"])) <= TIE_GAP * max(abs(float(z[k + "val"])), 1e-6 * val_scale), \
                (config, s, a_ref, float(z[k + "val"]))
            tie_sets.append(str(z[k + "name"]))
        np.testing.assert_allclose(out.set_values[s], float(z[k + "val"]), rtol=RTOL, atol=RTOL * 1e-6 * val_scale)
    assert out.set == int(z["selected_set"])      # the winning set's maximum is never a tie in these fixtures
    # (a reference run that sampled a large set counted its NaNs on the sample only: the oracle's full-grid count stands in)
    assert out.n_nan == sum(int((inputs if int(z.get(f"set{s}_sampled", 0)) else z)[f"set{s}_n_nan"]) for s in range(len(problems)))
    # The escape hatches, asserted: a different argmax is only ever accepted on a set whose expected acquisition is FLAT at its
    # maximum (top-two gap <= 1e-9: EI underflow plateaus far from every interventional row), and the floors of the relative
    # error apply to a bounded share of the compared candidates.
    flat_sets = [str(z[f"set{s}_name"]) for s in range(len(problems))
                 if min(float(z[f"set{s}_top2_gap"]), float(inputs[f"set{s}_top2_gap"])) <= TIE_GAP]
    assert set(tie_sets) <= set(flat_sets), (tie_sets, flat_sets)
    # (variance floor 1e-4 Kdiag: candidates next to an interventional row, 15 % of the toy graph's 1-D grids; mean floor:
    # zero crossings of mu)
    assert floored["var"] <= 0.2 * floored["kept"] and floored["mu"] <= 0.35 * floored["kept"], floored
    report = {"config": config, "expected": expected, "worst_rel_err": {k: float(f"{v:.3e}") for k, v in worst.items()},
              "argmax_ties_resolved_differently": tie_sets, "sets_with_flat_maximum": flat_sets, "floored_candidates": floored}
    print("PARITY", json.dumps(report))
