"""GPy.kern.src.stationary.Stationary: parameters and the scaled Euclidean distance (published GPy 1.10 behaviour)."""
import numpy as np
from paramz import Param, Parameterized
from paramz.transformations import Logexp

from ...util.linalg import tdot


class Stationary(Parameterized):
    _support_GPU = False

    def __init__(self, input_dim, variance, lengthscale, ARD, active_dims, name, useGPU=False):
        super().__init__(name=name)
        self.input_dim = int(input_dim)
        self.active_dims = active_dims
        self.useGPU = bool(useGPU) and self._support_GPU and False   # no GPU psi-statistics here
        self.ARD = ARD
        if not ARD:
            if lengthscale is None:
                lengthscale = np.ones(1)
            else:
                lengthscale = np.asarray(lengthscale)
                assert lengthscale.size == 1, "Only 1 lengthscale needed for non-ARD kernel"
        else:
            if lengthscale is not None:
                lengthscale = np.asarray(lengthscale)
                assert lengthscale.size in [1, input_dim], "Bad number of lengthscales"
                if lengthscale.size != input_dim:
                    lengthscale = np.ones(input_dim) * lengthscale
            else:
                lengthscale = np.ones(self.input_dim)
        self.lengthscale = Param('lengthscale', lengthscale, Logexp())
        self.variance = Param('variance', variance, Logexp())
        assert self.variance.size == 1
        self.link_parameters(self.variance, self.lengthscale)

    def _save_to_input_dict(self):
        return {"input_dim": self.input_dim, "variance": self.variance.tolist(), "lengthscale": self.lengthscale.tolist(),
                "ARD": self.ARD}

    def K_of_r(self, r):
        raise NotImplementedError

    def K(self, X, X2=None):
        r = self._scaled_dist(X, X2)
        return self.K_of_r(r)

    def Kdiag(self, X):
        ret = np.empty(X.shape[0])
        ret[:] = self.variance
        return ret

    def _unscaled_dist(self, X, X2=None):
        if X2 is None:
            Xsq = np.sum(np.square(X), 1)
            r2 = -2. * tdot(X) + (Xsq[:, None] + Xsq[None, :])
            r2[np.diag_indices(X.shape[0])] = 0.   # GPy forces the diagonal to zero in this branch only
            r2 = np.clip(r2, 0, np.inf)
            return np.sqrt(r2)
        X1sq = np.sum(np.square(X), 1)
        X2sq = np.sum(np.square(X2), 1)
        r2 = -2. * np.dot(X, X2.T) + (X1sq[:, None] + X2sq[None, :])
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)

    def _scaled_dist(self, X, X2=None):
        if self.ARD:
            if X2 is not None:
                X2 = X2 / self.lengthscale
            return self._unscaled_dist(X / self.lengthscale, X2)
        return self._unscaled_dist(X, X2) / self.lengthscale

    def update_gradients_diag(self, dL_dKdiag, X):
        raise NotImplementedError("gradients are outside the acquisition path")

    def update_gradients_full(self, dL_dK, X, X2=None):
        raise NotImplementedError("gradients are outside the acquisition path")

    def __getstate__(self):
        return dict(self.__dict__)

    def __setstate__(self, state):
        self.__dict__.update(state)
