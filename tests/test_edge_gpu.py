"""GPU parity at the edges of the shape space: the smallest and largest sizes the C ABI accepts, ragged sizes around the
128-row / 16-column padding, single-candidate grids, ranks that own no candidates of a set, and the ABI's own limits.
Same rules as test_parity_gpu.py (bit-exact integers, 1e-6 relative floats)."""
import numpy as np
import pytest

from helpers import RTOL, make_case, oracle_sweep, rel_err
from test_parity_gpu import _check_set, _engine

pytestmark = pytest.mark.gpu

EDGE_CASES = {
    "one_interventional_row": dict(seed=101, N=70, d=2, c=1, n=1, p=(9, 8)),
    "max_interventional_rows": dict(seed=102, N=150, d=3, c=1, n=128, p=(6, 5, 4), lo=-12.0, hi=12.0),   # 4 row passes of 32
    "single_candidate_grid": dict(seed=103, N=60, d=1, c=1, n=5, p=(1,)),
    "single_candidate_3d": dict(seed=104, N=60, d=3, c=0, n=5, p=(1, 1, 1)),
    "one_observation": dict(seed=105, N=1, d=1, c=1, n=4, p=(17,)),
    "two_observations_no_conditioning": dict(seed=106, N=2, d=2, c=0, n=4, p=(5, 5)),
    "n_obs_127": dict(seed=107, N=127, d=2, c=2, n=8, p=(12, 11)),
    "n_obs_129": dict(seed=108, N=129, d=2, c=2, n=8, p=(12, 11)),          # one row into the second 128-block
    "n_obs_385_three_blocks_plus_one": dict(seed=109, N=385, d=1, c=1, n=7, p=(130,)),
    "one_mc_sample": dict(seed=110, N=90, d=2, c=2, n=6, p=(7, 9), S_mc=1),
    "mc_samples_17": dict(seed=111, N=90, d=2, c=3, n=6, p=(7, 9), S_mc=17),  # one past the 16-column padding
    "max_conditioning_dims": dict(seed=112, N=100, d=1, c=8, n=6, p=(33,)),
    "four_intervened_dims": dict(seed=113, N=80, d=4, c=0, n=9, p=(4, 3, 5, 2)),
    "grid_of_129": dict(seed=114, N=64, d=1, c=1, n=6, p=(129,)),             # one candidate into the second tile
    "variable_cost_max_task": dict(seed=115, N=64, d=2, c=1, n=6, p=(10, 10), cost_variable=True),
}


@pytest.mark.parametrize("name", list(EDGE_CASES))
def test_edge_case(cuda_engine_ready, name):
    case = EDGE_CASES[name]
    task = "max" if name.endswith("max_task") else "min"
    kw, ora = make_case(**case)
    best = float(np.min(kw["y_int"]) if task == "min" else np.max(kw["y_int"]))
    ref = oracle_sweep(ora, best, task)
    eng = _engine([kw])
    out = eng.sweep(best, task)
    rep = _check_set(eng, 0, ref, ora, {})
    info = eng.fetch("fit_info", 0)
    assert info[1] == 0 and info[0] == ref["tries"]
    for k, v in rep.items():
        if not k.endswith("_pts"):
            assert v <= RTOL, f"{name} {k}: {v:.3e} (fit cond {rep.get('fit_cond_pts', 0):.1e})"
    assert out.index == ref["idx"] and out.n_nan == ref["n_nan"]
    np.testing.assert_allclose(out.value, ref["val"], rtol=RTOL)
    np.testing.assert_array_equal(out.x, ref["x"])


def test_rank_without_candidates(cuda_engine_ready):
    """Three ranks over two tiny sets: the partition leaves a rank with no candidates of a set (g_count = 0) and possibly
    with no set at all; every rank's stage calls must still run, and the slices must reproduce the unsplit arrays."""
    import torch
    from cbo_with_oop_b200.engine import SetProblem, SweepEngine
    cases = [make_case(seed=121, N=70, d=1, c=1, n=5, p=(100,)), make_case(seed=122, N=90, d=2, c=1, n=6, p=(13, 10))]
    best = float(min(np.min(k["y_int"]) for k, _ in cases))
    problems = [SetProblem(**k) for k, _ in cases]
    whole = SweepEngine(problems, keep=("acq",))
    whole.sweep(best, "min")
    full = [whole.fetch("acq", g) for g in range(2)]
    covered = [np.zeros(len(f), bool) for f in full]
    empty_seen = False
    for r in range(3):
        e = SweepEngine(problems, rank=r, world_size=3, keep=("acq",))
        e.build_tables(); e.prior_precompute(); e.prior_eval(1); e.posterior_fit(); e.prior_eval(0)
        e._sweep_local(best, "min")
        torch.cuda.synchronize()
        for g in range(2):
            gb, gc = e.slices[g]
            empty_seen |= gc == 0
            if gc:
                np.testing.assert_array_equal(e.fetch("acq", g), full[g][gb:gb + gc])
                assert not covered[g][gb:gb + gc].any()
                covered[g][gb:gb + gc] = True
    assert all(c.all() for c in covered) and empty_seen


def test_abi_limits_are_enforced(cuda_engine_ready):
    """Sizes past the limits of include/cbo_b200.h are refused with a message, not mis-computed."""
    from cbo_with_oop_b200._lib import CboError
    kw, _ = make_case(seed=131, N=40, d=2, c=1, n=5, p=(4, 4))
    kw["x_int"] = np.vstack([kw["x_int"]] * 26)[:129]     # 129 interventional rows > CBO_MAX_NINT
    kw["y_int"] = np.concatenate([kw["y_int"]] * 26)[:129]
    with pytest.raises((CboError, ValueError)):
        _engine([kw]).sweep(0.0, "min")
    kw, _ = make_case(seed=132, N=40, d=2, c=1, n=5, p=(4, 4))
    with pytest.raises(ValueError):
        _engine([kw]).sweep(0.0, "sideways")
