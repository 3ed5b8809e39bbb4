"""Convert the data shipped with the reference (/root/reference/data, read-only, only present in the build
container) into small pickle-free .npz fixtures under tests/golden/data/.  Run once; the outputs are committed.

    python tests/golden/import_reference_data.py [/root/reference]

Per experiment: the first 1000 observational rows (all columns), the optional true_observations table and the
interventional designs with their true causal effects (reference layout: row = [k, name_1..name_k, X, y],
SURVEY.md Appendix C).  src/DataLoader.py reads these when ./data/<experiment>/ (the reference layout) is absent.
"""
import os
import sys

import numpy as np
import pandas as pd

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
os.makedirs(out_dir, exist_ok=True)
for exp in ["toy_graph", "complete_graph", "coral_graph", "simplified_coral_graph"]:
    d = os.path.join(ref, "data", exp)
    obs = pd.read_pickle(os.path.join(d, "observations.pkl"))
    pack = {"columns": np.array(list(obs.columns)), "observations": np.asarray(obs, np.float64)[:1000]}
    tpath = os.path.join(d, "true_observations.pkl")
    if os.path.exists(tpath):
        t = pd.read_pickle(tpath)
        pack["true_columns"] = np.array(list(t.columns))
        pack["true_observations"] = np.asarray(t, np.float64)
    inter = np.load(os.path.join(d, "interventional_data.npy"), allow_pickle=True)
    pack["num_sets"] = np.array(len(inter))
    for j, row in enumerate(inter):
        k = int(row[0])
        pack[f"set{j}_names"] = np.array([str(v) for v in row[1:1 + k]])
        X = np.asarray(row[1 + k], np.float64)
        y = np.asarray([np.asarray(v, np.float64).reshape(-1)[0] for v in row[-1]]) if not isinstance(row[-1], np.ndarray) \
            else np.asarray(row[-1], np.float64)
        pack[f"set{j}_x"] = X.reshape(len(X), -1)
        pack[f"set{j}_y"] = np.asarray(y, np.float64).reshape(-1, 1)
    opt = os.path.join(d, "true_optimal_values.npy")
    if os.path.exists(opt):
        o = np.load(opt, allow_pickle=True)
        try:
            pack["true_optimal_y"] = np.array([float(np.asarray(v[-1]).reshape(-1)[0]) for v in o])
        except Exception:
            pass
    np.savez_compressed(os.path.join(out_dir, exp + ".npz"), **pack)
    print(exp, {k: v.shape for k, v in pack.items() if k in ("observations", "true_observations")}, int(pack["num_sets"]))
