"""Generate the golden parity fixtures tests/golden/golden_<config>.npz.

    python tests/golden/make_golden.py            (needs only this repository: reads tests/golden/data/*.npz)

For each BASELINE.json config that runs on shipped data (toy_graph, complete_graph, simplified_coral_graph; SURVEY.md
§8d rows 1-3) and for coral_graph with synthetic observations (row 4): build every exploration set's inputs exactly
as the agent does after its first observation (src.DataLoader + src.CBO + graph.fit_all_gaussian_processes with the
hyper-parameter fit frozen into the file), evaluate the ORACLE (oracle/cbo_oracle.py, factorised form, exact
distances) on the full 100-points-per-dimension grid and store
  * every input the CUDA path needs (so the GPU test does not depend on SciPy's optimiser),
  * per set: argmax index / value / coordinates, NaN count, jitter retries, L, alpha, prior at x_int,
  * per set: m, v, mu, var, ei, acq on a strided subsample of the grid (stride chosen so that <= 4096 points are kept).
The reference ships no golden vectors for this path: these are oracle outputs, "parity unpinned".
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import cbo_oracle as O  # noqa: E402
from src.DataLoader import DataLoader  # noqa: E402
from src.CBO import CBO  # noqa: E402
from src.utils_functions.utils import ObservationalGP  # noqa: E402
from cbo_with_oop_b200.obs_gp import fit_state, optimize_hyperparameters  # noqa: E402


def agent(experiment, n_obs, synthetic_obs=False, p=100):
    args = types.SimpleNamespace(exploration_set="MIS", initial_num_obs_samples=n_obs, num_interventions=10, type_cost=1,
                                 num_additional_observations=20, num_trials=3, name_index=0, seed=9, causal_prior=True,
                                 experiment=experiment, task="min", grid_points=p, device="cuda:0", num_sem_samples=1000)
    np.random.seed(args.seed)
    data = DataLoader(experiment, n_obs)
    if synthetic_obs:   # config 4: each column i.i.d. Normal(mean, std) of the shipped column, default_rng(1004)
        rng = np.random.default_rng(1004)
        cols = list(data.all_measurements.columns)
        mu, sd = data.all_measurements.mean().values, data.all_measurements.std().values
        import pandas as pd
        data.all_measurements = pd.DataFrame(mu + sd * rng.standard_normal((n_obs + 40, len(cols))), columns=cols)
        data.measurements = data.all_measurements[:n_obs]
        data.graph.measurements = {v: np.asarray(data.measurements[v]).reshape(-1, 1) for v in data.graph.var_names}
    return CBO(args, data, verbose=False)


def fit_gps(cbo):
    """Observational GPs with a bounded hyper-parameter fit (lengthscale >= half the column's std, variance <= 20):
    with the noise pinned to 1e-2 the unbounded optimum is a vanishing lengthscale, which makes every table exactly 0
    and would test nothing."""
    g = cbo.graph
    meas = {v: np.asarray(cbo.measurements[v], np.float64).reshape(-1, 1) for v in g.var_names}
    gps = {}
    for s in cbo.exploration_set:
        cols = g.prior_columns(s)
        name = g.get_gp_name(cols)
        if name in gps:
            continue
        x = np.hstack([meas[c] for c in cols])
        y = meas["Y"].reshape(-1)
        ard = bool(g.fit_parameters_for(cols)[3])
        floor = 0.5 * x.std(0) if ard else float(0.5 * x.std(0).max())
        s2, ls = optimize_hyperparameters(x, y, 1.0, 1.0, ard, 1e-2, min_lengthscale=floor, max_variance=20.0)
        alpha, kyinv = fit_state(x, y, s2, ls, 1e-2)
        gps[name] = ObservationalGP(x, y.reshape(-1, 1), s2, ls, 1e-2, alpha, kyinv, ard)
    return gps


def run(config, experiment, n_obs, synthetic=False, max_keep=4096, store_kyinv=True):
    cbo = agent(experiment, n_obs, synthetic)
    cbo.measurements = cbo.measurements  # first observe() is not needed: the GPs are fitted on the initial rows
    gps = fit_gps(cbo)
    cbo.do_calculus.gaussian_processes = gps
    best = float(cbo.current_best_solution())
    pack = {"config": np.array(config), "experiment": np.array(experiment), "num_sets": np.array(cbo.es_size),
            "best": np.array(best), "task": np.array("min")}
    vals = []
    for s in range(cbo.es_size):
        pr = cbo.do_calculus.set_problem(s)
        d = pr.d
        X = np.hstack([pr.x_obs_int, pr.x_obs_cond])
        ls = np.concatenate([pr.ls_int, pr.ls_cond])
        gp = dict(X=X, variance=pr.s2, lengthscale=ls, noise=pr.noise, alpha=pr.alpha_obs, Kyinv=pr.kyinv, form="diff")
        ref = O.sweep_set(gp, X, list(range(d)), pr.x_int, pr.y_int, pr.grid, best, "min", fix_costs=np.array([pr.cost_fix]),
                          variable_cost=pr.cost_variable, causal=True, prior="factorised", form="diff", precise_int=True)
        G = pr.g_total
        stride = max(1, -(-G // max_keep))
        keep = np.arange(0, G, stride)
        # every STORED value comes from the extended-precision prior (see do_prior_factorised): the interventional rows
        # (above) and the kept candidates + the argmax (here); the full-grid pass that located the argmax is float64
        pts = np.unique(np.append(keep, ref["idx"]))
        Xp = np.stack([pr.grid[a][np.unravel_index(pts, [len(t) for t in pr.grid])[a]] for a in range(d)], 1)
        mp, vp = O.do_prior_factorised(gp, ref["factors"], list(range(d)), Xp, precise=True)
        mup, varp = O.posterior_predict(ref["post"], Xp, mp, vp)
        eip = O.expected_improvement(mup, varp, best, "min")
        acqp = eip / O.point_cost(Xp, np.array([pr.cost_fix]), pr.cost_variable)
        for name, arr in (("mg", mp), ("vg", vp), ("mu", mup), ("var", varp), ("ei", eip), ("acq", acqp)):
            ref[name] = np.array(ref[name]); ref[name][pts] = arr
        ref["val"] = float(ref["acq"][ref["idx"]])
        k = f"set{s}_"
        # Ky^-1 (N x N per set) dominates the file size: the coral configs store the GP's training targets instead and
        # the test rebuilds (alpha, Ky^-1) with cbo_with_oop_b200.obs_gp.fit_state (deterministic LAPACK, agrees to ~1e-13)
        names = ["x_obs_int", "x_obs_cond", "alpha_obs", "ls_int", "ls_cond", "x_int", "y_int"] + (["kyinv"] if store_kyinv else [])
        for name in names:
            pack[k + name] = np.asarray(getattr(pr, name), np.float64)
        pack[k + "y_obs"] = np.asarray(gps[cbo.graph.get_gp_name(cbo.graph.prior_columns(cbo.exploration_set[s]))].Y, np.float64).reshape(-1)
        pack[k + "s2"] = np.array(pr.s2)
        pack[k + "cost_fix"] = np.array(pr.cost_fix)
        pack[k + "name"] = np.array(pr.name)
        pack[k + "grid_lo_hi_p"] = np.array([[t[0], t[-1], len(t)] for t in pr.grid])
        pack[k + "keep"] = keep
        for name in ["mg", "vg", "mu", "var", "ei", "acq"]:
            pack[k + name] = ref[name][keep]
        for name in ["mI", "vI", "L", "alpha"]:
            pack[k + name] = ref[name]
        pack[k + "idx"], pack[k + "val"], pack[k + "x"] = np.array(ref["idx"]), np.array(ref["val"]), ref["x"]
        pack[k + "n_nan"], pack[k + "tries"] = np.array(ref["n_nan"]), np.array(ref["tries"])
        srt = np.sort(ref["acq"][~np.isnan(ref["acq"])])
        pack[k + "top2_gap"] = np.array((srt[-1] - srt[-2]) / abs(srt[-1]) if len(srt) > 1 and srt[-1] != 0 else np.inf)
        vals.append(ref["val"])
        print(f"  {config} set {s} {pr.name:5s} N={X.shape[0]} D={X.shape[1]} G={G} idx={ref['idx']} val={ref['val']:.4e} "
              f"gap={float(pack[k + 'top2_gap']):.1e} nan={ref['n_nan']}", flush=True)
    sel, _ = O.select_set(vals)
    pack["selected_set"] = np.array(sel)
    out = os.path.join(HERE, f"golden_{config}.npz")
    np.savez_compressed(out, **pack)
    print(config, "->", out, os.path.getsize(out) // 1024, "KiB; selected set", sel)


if __name__ == "__main__":
    which = sys.argv[1:] or ["toy", "complete", "simplified_coral", "coral_synth"]
    if "toy" in which:
        run("toy", "toy_graph", 100)
    if "complete" in which:
        run("complete", "complete_graph", 100, max_keep=1024)
    if "simplified_coral" in which:
        run("simplified_coral", "simplified_coral_graph", 100, max_keep=512, store_kyinv=False)
    if "coral_synth" in which:
        run("coral_synth", "coral_graph", 200, synthetic=True, max_keep=512, store_kyinv=False)
