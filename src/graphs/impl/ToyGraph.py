"""X -> Z -> Y toy problem.  The reference ships data/toy_graph but no graph class; the
SEM below is the one the shipped data follow (X = eps, Z = exp(-X) + eps, Y = cos(Z) - exp(-Z/20) + eps; the true
E[Y|do(Z=z)] in interventional_data.npy equals cos z - exp(-z/20) to 9e-13)."""
from collections import OrderedDict

import numpy as np

from src.graphs.GraphInterface import GraphInterface


class ToyGraph(GraphInterface):
    adjustment = {"X": [], "Z": []}     # no back-door paths: E[Y|do(X)] = E[Y|X], E[Y|do(Z)] = E[Y|Z]
    cost_table = {1: ({"X": 1, "Z": 1}, False), 2: ({"X": 1, "Z": 10}, False),
                  3: ({"X": 1, "Z": 10}, True), 4: ({"X": 1, "Z": 1}, True)}

    def __init__(self, measurements, true_measurements=None):
        super().__init__(["X", "Z"])
        self.var_names = ["X", "Z", "Y"]
        self.measurements = {v: np.asarray(measurements[v], np.float64).reshape(-1, 1) for v in self.var_names}

    def define_sem(self):
        def f_x(epsilon, **kwargs):
            return epsilon[0]

        def f_z(epsilon, X, **kwargs):
            return np.exp(-X) + epsilon[1]

        def f_y(epsilon, Z, **kwargs):
            return np.cos(Z) - np.exp(-Z / 20.0) + epsilon[2]

        return OrderedDict([("X", f_x), ("Z", f_z), ("Y", f_y)])

    def device_sem(self):
        """define_sem() as a program for cbo_sem_eval (term = (source, func, coef, scale): coef * func(scale * source))."""
        return {"noise": ["e0", "e1", "e2"], "draw": self._draw_gaussian_noise(3),
                "nodes": [("X", 0.0, [("e0", "id", 1.0, 1.0)]),
                          ("Z", 0.0, [("X", "exp", 1.0, -1.0), ("e1", "id", 1.0, 1.0)]),
                          ("Y", 0.0, [("Z", "cos", 1.0, 1.0), ("Z", "exp", -1.0, -1.0 / 20.0), ("e2", "id", 1.0, 1.0)])]}

    @staticmethod
    def get_exploration_set(set_name):
        return [["X"], ["Z"]]

    @staticmethod
    def get_interventional_ranges():
        return OrderedDict([("X", [-5, 5]), ("Z", [-5, 20])])
