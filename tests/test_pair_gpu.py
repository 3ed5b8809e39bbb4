"""GPU parity of the small-N tensor-grid decomposition of the causal prior (csrc/prior_pair.cu: u^T M u regrouped over the
index pairs of M, one GEMM per grid plane) against the CPU oracle -- DoCalculus.update_do_function, DoCalculus.py:34-66.
The library takes that path only when a call holds at least one (scale row, tile) item per SM, so every case here is
shaped to reach that count (a long first grid dimension) and asserts, through cbo_prior_pair_items, that it did."""
import numpy as np
import pytest

from helpers import RTOL, make_case, oracle_sweep, rel_err

pytestmark = pytest.mark.gpu


def _engine(kws, **kw):
    from cbo_with_oop_b200.engine import SetProblem, SweepEngine
    return SweepEngine([SetProblem(**k) for k in kws], keep=("mu", "var", "ei", "acq"), **kw)


def _pair_items(eng):
    return int(eng.lib.cbo_prior_pair_items(eng.h_sets, len(eng.active), eng.num_sms))


def _check_prior(eng, g, ref, lo=None, hi=None):
    sl = slice(lo, hi)
    em = rel_err(eng.fetch("m", g), ref["mg"][sl], 1e-6).max()
    ev = rel_err(eng.fetch("v", g), ref["vg"][sl], 1e-6).max()
    assert em <= RTOL and ev <= RTOL, (g, em, ev)
    return em, ev


PAIR_CASES = [
    # one 3-D set: 150 scale rows of a 9 x 13 plane (2 x 2 live 8 x 8 blocks)
    [dict(seed=201, N=70, d=3, c=2, n=9, p=(150, 9, 13))],
    # chunked operands (130 rows -> 2 chunks of 72; 150 -> 2 of 80), a 2-D set riding along, a 1-D set and a non-causal
    # set in the same call (both stay on the general path)
    [dict(seed=202, N=33, d=3, c=1, n=8, p=(160, 130, 20)), dict(seed=203, N=100, d=2, c=2, n=10, p=(140, 150)),
     dict(seed=204, N=96, d=1, c=1, n=7, p=(50,)), dict(seed=205, N=64, d=2, c=1, n=6, p=(20, 20), causal=False)],
    # the largest observational set of this path (two 128-row blocks of M) and odd sizes; S_mc != N
    [dict(seed=206, N=256, d=3, c=1, n=12, p=(149, 8, 7), S_mc=90), dict(seed=207, N=17, d=3, c=0, n=5, p=(30, 3, 5))],
    # the shipped shape: 100-point dimensions (13 x 13 live blocks, the 7|6 x 4|3|3|3 warp split), ARD off
    [dict(seed=208, N=100, d=3, c=3, n=10, p=(148, 100, 100), ard=False)],
]


@pytest.mark.parametrize("specs", PAIR_CASES, ids=lambda s: "seed%d" % s[0]["seed"])
def test_pair_path_matches_oracle(cuda_engine_ready, specs):
    cases = [make_case(**s) for s in specs]
    best = float(min(np.min(k["y_int"]) for k, _ in cases))
    eng = _engine([k for k, _ in cases])
    assert _pair_items(eng) >= eng.num_sms, "the case was meant to take the pair-table path"
    out = eng.sweep(best, "min")
    refs = [oracle_sweep(o, best, "min") for _, o in cases]
    for g, ((_, ora), ref) in enumerate(zip(cases, refs)):
        if ora["causal"]:
            _check_prior(eng, g, ref)
        ei_scale = max(np.nanmax(np.abs(ref["ei"])), 1e-300)
        assert np.nanmax(rel_err(eng.fetch("acq", g), ref["acq"], 1e-6 * ei_scale)) <= RTOL
        assert out.set_indices[g] == ref["idx"], (g, out.set_indices[g], ref["idx"])
    s_ref = int(np.argmax([r["val"] for r in refs]))
    assert out.set == s_ref and out.index == refs[s_ref]["idx"]


def test_pair_path_on_a_slice_of_the_grid(cuda_engine_ready):
    """Two ranks' slices of one set (cut inside a grid plane): together they must reproduce the whole grid bit for bit."""
    spec = dict(seed=211, N=60, d=3, c=1, n=8, p=(321, 9, 13))
    kw, ora = make_case(**spec)
    best = float(np.min(kw["y_int"]))
    ref = oracle_sweep(ora, best, "min")
    whole = _engine([kw])
    whole.build_tables(), whole.prior_precompute(), whole.prior_eval(0)
    m_all, v_all = whole.fetch("m", 0), whole.fetch("v", 0)
    _check_prior(whole, 0, ref)
    plane, got_m, got_v = 9 * 13, [], []
    for rank in range(2):
        eng = _engine([kw], rank=rank, world_size=2)
        gb, gc = eng.slices[0]
        assert gc > 0 and (rank == 1 or (gb + gc) % plane != 0), "the cut was meant to fall inside a plane"
        assert _pair_items(eng) >= eng.num_sms
        eng.build_tables(), eng.prior_precompute(), eng.prior_eval(0)
        _check_prior(eng, 0, ref, gb, gb + gc)
        got_m.append(eng.fetch("m", 0)), got_v.append(eng.fetch("v", 0))
    np.testing.assert_array_equal(np.concatenate(got_m), m_all)
    np.testing.assert_array_equal(np.concatenate(got_v), v_all)


def test_pair_and_general_paths_agree(cuda_engine_ready):
    """The same set through both decompositions: a workspace without room for the pair tables forces the general kernel."""
    import ctypes as C
    import torch
    from cbo_with_oop_b200 import _lib
    kw, ora = make_case(seed=221, N=120, d=3, c=2, n=9, p=(150, 12, 10))
    eng = _engine([kw])
    eng.build_tables(), eng.prior_precompute(), eng.prior_eval(0)
    m_pair, v_pair = eng.fetch("m", 0), eng.fetch("v", 0)
    # the general layout for 8 CTAs is smaller than pair area + one slot: the library then runs the general kernel
    need_pair = eng.lib.cbo_prior_workspace_bytes(eng.h_sets, 1, 1)
    small = torch.empty((256 + 1024 * 2 * 128 * 8 + 8 * (4 * 2 * 128 * 8 + 128 * 128 * 8),), dtype=torch.uint8, device=eng.device)
    assert small.numel() < need_pair
    _lib.check(eng.lib.cbo_prior_eval(eng.h_sets, C.c_void_p(eng.d_sets.data_ptr()), 1, 0, C.c_void_p(small.data_ptr()),
                                      small.numel(), eng._stream()), "cbo_prior_eval")
    m_gen, v_gen = eng.fetch("m", 0), eng.fetch("v", 0)
    assert rel_err(m_pair, m_gen, 1e-6).max() <= 1e-9
    assert rel_err(v_pair, v_gen, 1e-6).max() <= 1e-9


def test_pair_path_first_launch_is_repeatable(cuda_engine_ready):
    """The first evaluation on a fresh workspace must equal the second bit for bit (the summation order is fixed).  Guards the
    second-level accumulator slots in shared memory: an earlier version read them before their first write and was wrong, on
    the first launch only, in the one warp that owned a 4 x 4 share of the 13 x 13 tile."""
    kw, _ = make_case(seed=231, N=100, d=3, c=2, n=8, p=(148, 100, 100), ard=False)
    for attempt in range(2):          # two fresh engines: fresh shared memory / workspace contents each time
        eng = _engine([kw])
        assert _pair_items(eng) >= eng.num_sms
        eng.build_tables(), eng.prior_precompute(), eng.prior_eval(0)
        m0, v0 = eng.fetch("m", 0).copy(), eng.fetch("v", 0).copy()
        for _ in range(2):
            eng.prior_eval(0)
            np.testing.assert_array_equal(eng.fetch("m", 0), m0)
            np.testing.assert_array_equal(eng.fetch("v", 0), v0)
        del eng
