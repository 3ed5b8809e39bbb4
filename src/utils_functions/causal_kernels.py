"""CausalRBF (reference: src/utils_functions/causal_kernels.py:45-79): k(x, x') = s2 exp(-r^2/2) + sqrt(v(x)) sqrt(v(x')).

In this build the expression is evaluated inside the CUDA kernels (csrc/posterior_fit.cu for K(X_I, X_I),
csrc/sweep.cu for K(X_I, x) and Kdiag(x)); this class only records the kernel's parameters and offers K / Kdiag as
host-side inspection helpers.  The psi-statistics / GPy-GPU boilerplate of the reference (:93-155) has no counterpart."""
import numpy as np


class CausalRBF:
    def __init__(self, input_dim, variance_adjustment, variance=1.0, lengthscale=1.0, rescale_variance=1.0, ARD=False,
                 active_dims=None, name="rbf", useGPU=False, inv_l=False):
        self.input_dim = input_dim
        self.variance_adjustment = variance_adjustment
        self.variance = float(variance)
        self.lengthscale = 1.0 if lengthscale is None else float(np.asarray(lengthscale).reshape(-1)[0])
        self.rescale_variance = rescale_variance
        self.ARD = ARD
        self.name = name

    def _r2(self, X, X2):
        d = (X[:, None, :] - X2[None, :, :]) / self.lengthscale
        return np.sum(d * d, axis=2)

    def K(self, X, X2=None):
        X = np.atleast_2d(np.asarray(X, np.float64))
        X2 = X if X2 is None else np.atleast_2d(np.asarray(X2, np.float64))
        vx = np.asarray(self.variance_adjustment(X), np.float64).reshape(-1, 1)
        vx2 = np.asarray(self.variance_adjustment(X2), np.float64).reshape(-1, 1)
        return self.variance * np.exp(-0.5 * self._r2(X, X2)) + np.sqrt(vx) @ np.sqrt(vx2).T

    def Kdiag(self, X):
        X = np.atleast_2d(np.asarray(X, np.float64))
        return self.variance + np.asarray(self.variance_adjustment(X), np.float64).reshape(-1)

    def K_of_r(self, r):
        return self.variance * np.exp(-0.5 * r ** 2)
