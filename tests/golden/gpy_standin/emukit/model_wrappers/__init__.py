from ..core.interfaces import IDifferentiable, IModel


class GPyModelWrapper(IModel, IDifferentiable):
    """emukit.model_wrappers.GPyModelWrapper: forwards to the wrapped GPy model."""

    def __init__(self, gpy_model, n_restarts=1):
        self.model = gpy_model
        self.n_restarts = n_restarts

    def predict(self, X):
        return self.model.predict(X)

    def set_data(self, X, Y):
        self.model.set_XY(X, Y)

    def optimize(self, verbose=False):
        self.model.optimize()

    @property
    def X(self):
        return self.model.X

    @property
    def Y(self):
        return self.model.Y
