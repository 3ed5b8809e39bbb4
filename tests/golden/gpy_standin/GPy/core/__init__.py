"""GPy.core: Param, Mapping (what GaussianProcessFactory.py:1 and causal_kernels.py:4 import)."""
from paramz import Param, Parameterized  # noqa: F401


class Mapping(Parameterized):
    """GPy.core.Mapping(input_dim, output_dim, name): base class of mean functions; the reference overwrites ``f``
    and ``update_gradients`` on an instance (GaussianProcessFactory.py:66-68)."""

    def __init__(self, input_dim, output_dim, name="mapping"):
        super().__init__(name=name)
        self.input_dim = input_dim
        self.output_dim = output_dim

    def f(self, X):
        raise NotImplementedError

    def update_gradients(self, dL_dF, X):
        raise NotImplementedError
