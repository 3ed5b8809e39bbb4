"""Debug aid: the mixed-geometry pair-path case (tests/test_pair_gpu.py seed202) repeated on fresh engines; prints where runs differ."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import make_case
from cbo_with_oop_b200.engine import SetProblem, SweepEngine
specs = [dict(seed=202, N=33, d=3, c=1, n=8, p=(160, 130, 20)), dict(seed=203, N=100, d=2, c=2, n=10, p=(140, 150))]
cases = [make_case(**s) for s in specs]
ref = None
for rep in range(int(os.environ.get("REPS", "12"))):
    eng = SweepEngine([SetProblem(**k) for k, _ in cases])
    eng.build_tables(); eng.prior_precompute(); eng.prior_eval(0)
    cur = [(eng.fetch("m", g).copy(), eng.fetch("v", g).copy()) for g in range(2)]
    if ref is None:
        eng.prior_eval(0)
        ref = [(eng.fetch("m", g).copy(), eng.fetch("v", g).copy()) for g in range(2)]
    for g in range(2):
        for name, a, b in (("m", cur[g][0], ref[g][0]), ("v", cur[g][1], ref[g][1])):
            bad = np.nonzero(a != b)[0]
            if bad.size:
                idx = np.unravel_index(bad, specs[g]["p"])
                print("rep", rep, "set", g, name, "bad", bad.size, [(int(np.min(i)), int(np.max(i)), np.unique(i).size) for i in idx],
                      "max rel", float(np.max(np.abs(a - b)[bad] / np.maximum(np.abs(b[bad]), 1e-300))))
    del eng
print("done")
