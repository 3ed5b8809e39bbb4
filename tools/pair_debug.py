"""Debug aid: the same set through the pair-table and the general decomposition; prints where they differ."""
import ctypes as C, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import make_case
from cbo_with_oop_b200 import _lib
from cbo_with_oop_b200.engine import SetProblem, SweepEngine
p = (148, 100, 100)
kw, ora = make_case(seed=208, N=100, d=3, c=3, n=10, p=p, ard=False)
eng = SweepEngine([SetProblem(**kw)])
eng.build_tables(); eng.prior_precompute()
res = []
for rep in range(int(os.environ.get("REPS", "8"))):
    eng.prior_eval(0)
    res.append((eng.fetch("m", 0).copy(), eng.fetch("v", 0).copy()))
small = torch.empty((256 + 1024 * 2 * 128 * 8 + 8 * (4 * 2 * 128 * 8 + 128 * 128 * 8),), dtype=torch.uint8, device=eng.device)
_lib.check(eng.lib.cbo_prior_eval(eng.h_sets, C.c_void_p(eng.d_sets.data_ptr()), 1, 0, C.c_void_p(small.data_ptr()), small.numel(), eng._stream()), "x")
mg, vg = eng.fetch("m", 0), eng.fetch("v", 0)
for rep, (m, v) in enumerate(res):
    for name, a, b in (("m", m, mg), ("v", v, vg)):
        bad = np.nonzero(np.abs(a - b) > 1e-7 * (1 + np.abs(b)))[0]
        print(rep, name, "bad", bad.size)
        if bad.size:
            i0, i1, i2 = np.unravel_index(bad, p)
            print("  i0:", np.unique(i0)[:20], "n", np.unique(i0).size)
            print("  row blocks:", np.unique(i1 // 8), "col blocks:", np.unique(i2 // 8))
            print("  items mod 148:", np.unique(i0 % 148)[:20])
            print("  sample diffs:", (a - b)[bad[:5]], b[bad[:5]])
print("repeatable:", all(np.array_equal(res[0][1], r[1]) for r in res))
# bitwise comparison of rep 0 against rep 1 (same decomposition, deterministic order): where do they differ at all?
for name, k in (("m", 0), ("v", 1)):
    a, b = res[0][k], res[1][k]
    bad = np.nonzero(a != b)[0]
    print("bitwise", name, bad.size)
    if bad.size:
        i0, i1, i2 = np.unravel_index(bad, p)
        print("  i0", np.unique(i0), "rows", np.unique(i1)[[0, -1]], np.unique(i1).size, "cols", np.unique(i2)[[0, -1]], np.unique(i2).size)
        rel = np.abs(a - b)[bad] / np.abs(b[bad])
        print("  rel err min/max", rel.min(), rel.max())
        # error as a function of column for the first bad row
        r0 = np.unique(i1)[0]
        sel = bad[(i1 == r0)]
        print("  row", r0, "cols", np.unravel_index(sel, p)[2][:40])
