"""CPU checks of the int8-slice emulation behind DESIGN.md §8's go / no-go (tools/ozaki_accuracy.py): the signed 7-bit
digits stay inside int8 and reconstruct their input, enough slices reproduce the float64 quadratic form, and too few do
not -- so the error table in profiles/r02_ozaki_accuracy.json measures the scheme, not the emulation."""
import importlib.util
import os

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
spec = importlib.util.spec_from_file_location("ozaki_accuracy", os.path.join(ROOT, "tools", "ozaki_accuracy.py"))
oz = importlib.util.module_from_spec(spec)
spec.loader.exec_module(oz)


def test_digits_fit_int8_and_reconstruct():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((7, 50)) * np.exp(rng.uniform(-30, 5, (7, 1)))
    A[3, :] = 0.0                      # an all-zero row keeps exponent 0 and zero digits
    for s in (1, 4, 8):
        d, e = oz.slices(A, s)
        assert np.abs(d).max() <= 64
        rec = sum(d[i].astype(np.float64) * 2.0 ** (-oz.BITS * (i + 1)) for i in range(s)) * np.exp2(e)[:, None]
        bound = np.exp2(e)[:, None] * 2.0 ** (-oz.BITS * s - 1)          # half a unit of the last digit, per row
        assert np.all(np.abs(rec - A) <= bound * (1 + 1e-12))


def test_quadratic_form_converges_with_slices():
    rng = np.random.default_rng(1)
    N, G = 40, 9
    B = rng.standard_normal((N, N))
    M = B @ B.T / N                       # symmetric
    U = np.exp(-rng.uniform(0, 6, (G, N)))
    ref = np.einsum("gk,gk->g", U.astype(np.longdouble) @ M.astype(np.longdouble), U.astype(np.longdouble)).astype(np.float64)
    err = {}
    for s in (3, 6, 9):
        q, peak = oz.quad_form_sliced(U, M, s)
        assert peak < 2 ** 31             # the int32 accumulator of the tensor core would not overflow
        err[s] = np.abs(q - ref).max() / np.abs(ref).max()
    assert err[9] < 1e-14 and err[6] < 1e-9 and err[3] > 1e-7 and err[3] > err[6] > err[9]
