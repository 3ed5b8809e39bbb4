"""Run the REFERENCE's own modules (unmodified, imported from /root/reference) on the inputs frozen in
tests/golden/golden_<config>.npz and store what they return -> tests/golden/reference_run_<config>.npz.

    python tests/golden/make_reference_golden.py [toy complete simplified_coral]        (build container only: needs /root/reference)

The reference's third-party dependencies (GPy, emukit, paramz) are absent from this image; tests/golden/gpy_standin/
supplies the API surface the reference imports (see its README for exactly what is and is not pinned by this).
This script must not import this repository's `src` package (same name as the reference's) nor `oracle/`.

Reference code executed per exploration set (file:line in /root/reference):
  * utils.py:40-45      fit_gaussian_process             observational GP (hyper-parameters = the frozen inputs)
  * DoCalculus.py:68-89 compute_do / get_intervened_inputs  intervened design + gp.predict, called from a closure that
                        mirrors update_do_function (:34-66): memo by str(value), then np.mean over the samples.  The
                        literal line :59 assigns an (N,1) array into a scalar slot and raises for N > 1 (SURVEY.md
                        App. B #6); the closure applies the np.mean of line :60 first, which is the evident intent.
                        The GP / dependency lookups of :45-47,:58 are replaced by the explicit column table (App. B #3-5).
  * GaussianProcessFactory.py:24-46,63-73  create(CAUSAL_GP, ..., emukit_wrapper=True)   with causal_kernels.py CausalRBF
  * causal_acquisition_functions.py:27-43  CausalExpectedImprovement.evaluate
  * cost_functions.py:11-17 Cost.evaluate, GraphInterface.py:46-50 cost, utils.py:34 the quotient
  * utils.py:8-26       find_current_global
  * CBO.py:269-277      select_next_intervention
The candidate set is the full 100-points-per-dimension grid (C order, last dimension fastest) instead of the
reference's 100 random anchors + L-BFGS; within a set the first maximum is taken.
"""
import os
import sys
import types
from collections import OrderedDict
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CBO_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "gpy_standin"), REF]
assert "src" not in sys.modules

from src.DoCalculus import DoCalculus                                    # noqa: E402  (the reference's)
from src.GaussianProcessFactory import GaussianProcessFactory, GaussianProcessType  # noqa: E402
from src.utils_functions import (CausalExpectedImprovement, Cost, find_current_global,  # noqa: E402
                                 fit_gaussian_process)
from src.graphs.GraphInterface import GraphInterface                    # noqa: E402
from src.CBO import CBO                                                  # noqa: E402
import src as _ref_src                                                   # noqa: E402

assert os.path.realpath(_ref_src.__file__).startswith(os.path.realpath(REF)), "the reference's src must be the one imported"


def set_variables(name, d):
    """'BD' -> ['B', 'D'] (every variable of the shipped graphs is one letter except coral's, not used here)."""
    assert len(name) == d
    return list(name)


def do_closure(do, measurements, gp, input_vars, intervention_vars, index, memo):
    """DoCalculus.update_do_function (DoCalculus.py:34-66) with compute_do called as at :59 and the mean of :60."""
    def f(values):
        out = np.zeros((values.shape[0], 1))
        for i, value in enumerate(values):
            key = str(value)
            if key in memo:
                out[i] = memo[key]
            else:
                out[i] = np.mean(do.compute_do(measurements, gp, value, input_vars, intervention_vars)[index])
            if key not in memo:
                memo[key] = out[i]
        return np.float64(out)
    return f


def run(config, max_keep, max_full=20000, sample=4096):
    """Sets with at most `max_full` candidates are evaluated on the whole grid.  Larger ones (the 1e6-candidate 3-D sets of
    the coral family: about 1 ms per candidate through the reference's per-candidate gp.predict) are evaluated on a seeded
    sample of `sample` candidates plus the 3^d neighbourhood of the oracle's argmax (golden_<config>.npz), flagged
    `sampled`: their `idx` / `val` are the maximum over that subset and `n_nan` counts the subset only."""
    z = np.load(os.path.join(HERE, f"golden_{config}.npz"))
    S = int(z["num_sets"])
    task = str(z["task"])
    names = [str(z[f"set{s}_name"]) for s in range(S)]
    # best-so-far through the reference's own helper
    current_y = {names[s]: list(z[f"set{s}_y_int"].reshape(-1)) for s in range(S)}
    best = find_current_global(current_y, names, task)
    pack = {"config": np.array(config), "best": np.array(float(best)), "task": np.array(task), "num_sets": np.array(S)}
    do = DoCalculus(types.SimpleNamespace())
    ys = []
    for s in range(S):
        k = f"set{s}_"
        xo_i, xo_c = z[k + "x_obs_int"], z[k + "x_obs_cond"]
        N, d = xo_i.shape
        c = xo_c.shape[1] if xo_c.size else 0
        iv = set_variables(names[s], d)
        cv = [f"cond{j}" for j in range(c)]
        input_vars = iv + cv
        measurements = {v: xo_i[:, j] for j, v in enumerate(iv)}
        measurements.update({v: xo_c[:, j] for j, v in enumerate(cv)})
        x = np.hstack([xo_i] + ([xo_c] if c else []))
        ls = np.concatenate([z[k + "ls_int"].reshape(-1), z[k + "ls_cond"].reshape(-1)])
        ard = x.shape[1] > 1
        gp = fit_gaussian_process(x, z[k + "y_obs"].reshape(-1, 1), [ls if ard else ls[:1], float(z[k + "s2"]), 1.0, ard])
        assert float(gp.likelihood.variance[0]) == 1e-2
        mean_fn = do_closure(do, measurements, gp, input_vars, iv, 0, {})
        var_fn = do_closure(do, measurements, gp, input_vars, iv, 1, {})
        XI, yI = z[k + "x_int"], z[k + "y_int"].reshape(-1, 1)
        model = GaussianProcessFactory.create(GaussianProcessType.CAUSAL_GP, XI, yI, [mean_fn, var_fn], emukit_wrapper=True)
        costs = OrderedDict((v, partial(GraphInterface.cost, float(z[k + "cost_fix"]) / d, False)) for v in iv)
        ei_acq = CausalExpectedImprovement(best, task, model)
        acquisition = ei_acq / Cost(costs, iv)
        tables = [np.linspace(lo, hi, int(p)) for lo, hi, p in z[k + "grid_lo_hi_p"]]
        shape = [len(t) for t in tables]
        G_full = int(np.prod(shape))
        sampled = G_full > max_full
        if sampled:
            rng = np.random.default_rng(2000 + s)
            centre = np.unravel_index(int(z[k + "idx"]), shape)
            nb = np.stack(np.meshgrid(*[np.clip(np.arange(c_ - 1, c_ + 2), 0, p_ - 1) for c_, p_ in zip(centre, shape)],
                                      indexing="ij"), axis=-1).reshape(-1, len(shape))
            flat = np.unique(np.concatenate([rng.choice(G_full, size=sample, replace=False),
                                             np.ravel_multi_index(nb.T, shape)]))
        else:
            flat = np.arange(G_full)
        ii = np.unravel_index(flat, shape)
        Xg = np.stack([tables[a][ii[a]] for a in range(len(shape))], axis=1)
        G = Xg.shape[0]
        acq = np.empty(G)
        mu, var, ei, mg, vg = (np.empty(G) for _ in range(5))
        for a in range(0, G, 1000):        # candidate batches; Cost.evaluate's batch quirk (App. B #12) is inert for fixed costs
            xb = Xg[a:a + 1000]
            acq[a:a + 1000] = acquisition.evaluate(xb)[:, 0]
            m_, v_ = model.predict(xb)
            mu[a:a + 1000], var[a:a + 1000] = m_[:, 0], v_[:, 0]
            ei[a:a + 1000] = ei_acq.evaluate(xb)[:, 0]
            mg[a:a + 1000], vg[a:a + 1000] = mean_fn(xb)[:, 0], var_fn(xb)[:, 0]
        nan = np.isnan(acq)
        idx = int(np.argmax(np.where(nan, -np.inf, acq)))
        ys.append(np.array([[acq[idx]]]))
        keep = np.arange(G) if sampled else np.arange(0, G, max(1, -(-G // max_keep)))
        pack[k + "sampled"] = np.array(int(sampled))
        post = model.model.posterior
        pack.update({k + "keep": flat[keep], k + "idx": np.array(int(flat[idx])), k + "val": np.array(acq[idx]), k + "x": Xg[idx],
                     k + "n_nan": np.array(int(nan.sum())), k + "tries": np.array(int(post.jitter_tries)),
                     k + "mI": mean_fn(XI)[:, 0], k + "vI": var_fn(XI)[:, 0], k + "L": post.woodbury_chol,
                     k + "alpha": post.woodbury_vector[:, 0], k + "name": np.array(names[s])})
        for nm, arr in (("mg", mg), ("vg", vg), ("mu", mu), ("var", var), ("ei", ei), ("acq", acq)):
            pack[k + nm] = arr[keep]
        srt = np.sort(acq[~nan])
        pack[k + "top2_gap"] = np.array((srt[-1] - srt[-2]) / abs(srt[-1]) if len(srt) > 1 and srt[-1] != 0 else np.inf)
        print(f"  {config} set {s} {names[s]:3s} N={N} D={x.shape[1]} G={G}{' (sampled)' if sampled else ''} idx={int(flat[idx])} val={acq[idx]:.6e} "
              f"tries={post.jitter_tries} nan={int(nan.sum())}", flush=True)
    agent = types.SimpleNamespace(monitor=types.SimpleNamespace(last_intervention=None),
                                  exploration_set=[set_variables(n, len(n)) for n in names])
    chosen_set, chosen = CBO.select_next_intervention(agent, ys)
    pack["selected_set"] = np.array(int(chosen))
    pack["selected_names"] = np.array("".join(chosen_set))
    out = os.path.join(HERE, f"reference_run_{config}.npz")
    np.savez_compressed(out, **pack)
    print(config, "->", out, os.path.getsize(out) // 1024, "KiB; selected set", int(chosen), chosen_set)


if __name__ == "__main__":
    which = sys.argv[1:] or ["toy", "complete"]
    if "toy" in which:
        run("toy", 4096)
    if "complete" in which:
        run("complete", 1024)
    if "simplified_coral" in which:      # 25 sets: 1-D and 2-D sets in full, 3-D sets sampled (about ten minutes)
        run("simplified_coral", 1024)
