import numpy as np

from .stationary import Stationary


class RBF(Stationary):
    """GPy.kern.RBF: k(r) = variance * exp(-0.5 r^2)."""

    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False, active_dims=None, name='rbf', useGPU=False,
                 inv_l=False):
        super().__init__(input_dim, variance, lengthscale, ARD, active_dims, name, useGPU=useGPU)

    def K_of_r(self, r):
        return self.variance * np.exp(-0.5 * r ** 2)
