"""Accuracy half of the int8-slice (Ozaki scheme) question of DESIGN.md §8 -- CPU only, exact integer emulation.

K1b's quadratic form  v(x) = s2 + noise - u^T M u  is evaluated on the golden fixtures the way an int8 tensor-core kernel
would: every row of U (candidates) and every row of the symmetric M is scaled by a power of two (its largest magnitude -> [0.25, 0.5))
and cut into `s` signed 7-bit slices; the slice products A_i B_j^T are exact in int32 (the tensor core's accumulator; the sums are formed here in
int64 and checked against the int32 range), only the pairs with i + j <= s + 1 are formed (s (s + 1) / 2 products, the scheme's
cost), equal-weight pairs are summed exactly (the 2 s - 1 accumulator groups), and the groups are recombined in float64 from
the lightest weight up; the row dot with u and the subtraction are float64, as in the epilogue of K1b.  The result is
compared with the extended-precision oracle (do_prior_factorised(precise=True)) and with the plain float64 evaluation (what
the DMMA kernel delivers), on the candidates each fixture keeps.

    python tools/ozaki_accuracy.py [coral_synth simplified_coral ...] > profiles/r02_ozaki_accuracy.json
The throughput half is tools/i8_umma_peak.cu (profiles/r02_i8_umma_peak.json)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from cbo_with_oop_b200.obs_gp import fit_state      # host input preparation (Ky^-1 of the fixture's observational GP)
from oracle import cbo_oracle as O                    # the checker (this is a test-side study, not a product path)

BITS = 7


def slices(A, s):
    """Rows of A -> (ints (s, rows, cols) with |digit| <= 64, exponent per row): A ~ 2^e sum_i digit_i 2^(-7 (i + 1))."""
    amax = np.abs(A).max(axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.where(amax > 0, amax, 1.0))) + 2, 0).astype(np.int64)      # |A| / 2^e in [0.25, 0.5): every digit within +-64
    r = A / np.exp2(e)[:, None]
    out = np.zeros((s,) + A.shape, np.int64)
    for i in range(s):            # round-to-nearest signed digits: the remainder stays within half a unit of the next digit
        d = np.rint(r * 2.0 ** BITS)
        out[i] = d.astype(np.int64)
        r = r * 2.0 ** BITS - d   # exact in float64 (a scaling by 2^7 and a subtraction of the rounded value)
    return out, e


def quad_form_sliced(U, M, s):
    """u^T M u per row of U through the slice products; returns (values, largest |group sum| seen)."""
    Us, eu = slices(U, s)
    Ms, em = slices(M, s)            # rows of M (M is symmetric: T = U M^T)
    G, N = U.shape
    T = np.zeros((G, N))
    peak = 0
    for g in range(2 * s - 2, -1, -1):          # weight 2^(-7 (g + 2)), lightest first; pairs (i, j), i + j = g, kept if g <= s - 1
        if g > s - 1:
            continue
        acc = np.zeros((G, N), np.int64)
        for i in range(0, g + 1):
            j = g - i
            if i < s and j < s:
                acc += Us[i] @ Ms[j].T
        peak = max(peak, int(np.abs(acc).max()))
        T += acc.astype(np.float64) * 2.0 ** (-BITS * (g + 2))
    T *= np.exp2(eu)[:, None] * np.exp2(em)[None, :]
    return np.einsum("gk,gk->g", T, U), peak


def study(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", f"golden_{name}.npz"), allow_pickle=True)
    rows = []
    for sidx in range(int(z["num_sets"])):
        k = f"set{sidx}_"
        X = np.hstack([z[k + "x_obs_int"], z[k + "x_obs_cond"]])
        d = z[k + "x_obs_int"].shape[1]
        ls = np.concatenate([z[k + "ls_int"], z[k + "ls_cond"]])
        s2 = float(z[k + "s2"])
        kyinv = z[k + "kyinv"] if k + "kyinv" in z else fit_state(X, z[k + "y_obs"], s2, ls, 1e-2)[1]
        gp = dict(X=X, variance=s2, lengthscale=ls, noise=1e-2, alpha=z[k + "alpha_obs"], Kyinv=kyinv, form="diff")
        cols = list(range(d))
        f = O.prior_factors(gp, X, cols)
        grid = [np.linspace(lo, hi, int(p)) for lo, hi, p in z[k + "grid_lo_hi_p"]]
        ii = np.unravel_index(z[k + "keep"], [len(t) for t in grid])
        Xg = np.stack([grid[a][ii[a]] for a in range(d)], axis=1)[:256]
        _, v_ref = O.do_prior_factorised(gp, f, cols, Xg, precise=True)
        U = O.intervened_u(gp, cols, Xg)
        q64 = np.einsum("gk,gk->g", U @ f["M"], U)
        v64 = (s2 + 1e-2) - q64
        cancel = float((np.einsum("gk,gk->g", np.abs(U) @ np.abs(f["M"]), np.abs(U)) / np.maximum(np.abs(v_ref), 1e-300)).max())
        row = {"set": str(z[k + "name"]), "n_obs": int(X.shape[0]), "candidates": int(Xg.shape[0]), "cancellation_max": cancel,
               "fp64_rel_err_v": float((np.abs(v64 - v_ref) / np.maximum(np.abs(v_ref), 1e-6)).max())}
        for s in (5, 6, 7, 8, 9):
            q, peak = quad_form_sliced(U, f["M"], s)
            v = (s2 + 1e-2) - q
            row[f"slices_{s}"] = {"products": s * (s + 1) // 2, "rel_err_v": float((np.abs(v - v_ref) / np.maximum(np.abs(v_ref), 1e-6)).max()),
                                  "int32_headroom_bits": float(31 - np.log2(max(peak, 1)))}
        rows.append(row)
    return rows


if __name__ == "__main__":
    names = sys.argv[1:] or ["coral_synth", "simplified_coral", "complete"]
    out = {"what": "relative error of v = s2 + noise - u^T M u against the extended-precision oracle: float64 (DMMA kernel) vs exact int8 "
                   "slice products with s signed 7-bit slices per operand row and the i + j <= s + 1 truncation (s (s + 1) / 2 products)",
           "bits_per_slice": BITS, "configs": {}}
    for n in names:
        rows = study(n)
        worst = {key: max(r[key] if not isinstance(r[key], dict) else r[key]["rel_err_v"] for r in rows)
                 for key in ["fp64_rel_err_v"] + [f"slices_{s}" for s in (5, 6, 7, 8, 9)]}
        out["configs"][n] = {"worst_over_sets": worst, "max_cancellation": max(r["cancellation_max"] for r in rows), "sets": rows}
    print(json.dumps(out, indent=1))
