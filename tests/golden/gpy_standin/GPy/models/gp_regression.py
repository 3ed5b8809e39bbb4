"""GPy.models.GPRegression with exact Gaussian inference, restated from the published GPy 1.10 algorithm
(ExactGaussianInference.inference, PosteriorExact._raw_predict, GP.predict, Gaussian.predictive_values)."""
import numpy as np
from paramz import Param, Parameterized
from paramz.transformations import Logexp

from ..kern import RBF
from ..util.linalg import dpotrs, dtrtrs, jitchol, pdinv


class Gaussian(Parameterized):
    def __init__(self, variance=1., name='Gaussian_noise'):
        super().__init__(name=name)
        self.variance = Param('variance', variance, Logexp())
        self.link_parameter(self.variance)

    def predictive_values(self, mu, var, full_cov=False, Y_metadata=None):
        return mu, var + self.variance       # include_likelihood=True: the noise is added, no +1e-8 here


class _Posterior:
    def __init__(self, woodbury_chol, woodbury_vector, K, woodbury_inv, jitter_tries):
        self.woodbury_chol, self.woodbury_vector, self.K = woodbury_chol, woodbury_vector, K
        self.woodbury_inv, self.jitter_tries = woodbury_inv, jitter_tries

    def _raw_predict(self, kern, Xnew, pred_var, full_cov=False):
        Kx = kern.K(pred_var, Xnew)
        mu = np.dot(Kx.T, self.woodbury_vector)
        if mu.ndim == 1:
            mu = mu.reshape(-1, 1)
        Kxx = kern.Kdiag(Xnew)
        tmp = dtrtrs(self.woodbury_chol, Kx)[0]
        var = (Kxx - np.square(tmp).sum(0))[:, None]
        return mu, var


class GPRegression(Parameterized):
    def __init__(self, X, Y, kernel=None, Y_metadata=None, normalizer=None, noise_var=1., mean_function=None):
        super().__init__(name='GP regression')
        assert X.ndim == 2 and Y.ndim == 2 and normalizer is None
        self.X, self.Y = np.array(X, dtype=np.float64), np.array(Y, dtype=np.float64)
        self.kern = RBF(X.shape[1]) if kernel is None else kernel
        self.likelihood = Gaussian(variance=noise_var)
        self.mean_function = mean_function
        self._posterior = None

    # -- inference ------------------------------------------------------------------------------------
    def _infer(self):
        m = 0 if self.mean_function is None else self.mean_function.f(self.X)
        YYT_factor = self.Y - m
        K = self.kern.K(self.X)
        Ky = K.copy()
        Ky[np.diag_indices(Ky.shape[0])] += float(self.likelihood.variance[0]) + 1e-8
        Wi, LW, LWi, W_logdet = pdinv(Ky)
        alpha, _ = dpotrs(LW, YYT_factor, lower=1)
        self._posterior = _Posterior(LW, alpha, K, Wi, jitchol.last_tries)

    @property
    def posterior(self):
        # Parameters are plain arrays here (no observer pattern): infer lazily, so a `.fix(value)` issued after
        # construction (utils.py:43) is honoured exactly as GPy's parameters_changed() would.
        if self._posterior is None:
            self._infer()
        return self._posterior

    def parameters_changed(self):
        self._posterior = None

    def set_XY(self, X, Y):
        self.X, self.Y = np.array(X, dtype=np.float64), np.array(Y, dtype=np.float64)
        self._posterior = None

    def optimize(self, *a, **kw):
        """Hyper-parameters are inputs of the path: nothing to do."""
        self._posterior = None
        return self

    # -- prediction -----------------------------------------------------------------------------------
    def _raw_predict(self, Xnew, full_cov=False, kern=None):
        mu, var = self.posterior._raw_predict(self.kern if kern is None else kern, Xnew, self.X, full_cov)
        if self.mean_function is not None:
            mu = mu + self.mean_function.f(Xnew)
        return mu, var

    def predict(self, Xnew, full_cov=False, Y_metadata=None, kern=None, likelihood=None, include_likelihood=True):
        mean, var = self._raw_predict(Xnew, full_cov=full_cov, kern=kern)
        if include_likelihood:
            mean, var = self.likelihood.predictive_values(mean, var, full_cov, Y_metadata=Y_metadata)
        return mean, var
