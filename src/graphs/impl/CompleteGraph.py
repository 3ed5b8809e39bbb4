"""Synthetic graph of the CBO paper (reference: src/graphs/impl/CompleteGraph.py): U1, U2 unobserved confounders,
manipulative variables B, D, E, target Y."""
from collections import OrderedDict

import numpy as np

from src.graphs.GraphInterface import GraphInterface


class CompleteGraph(GraphInterface):
    # conditioning columns of the observational GP per manipulative variable; the union over a set reproduces the
    # reference's fit_dependencies rows ["B"], ["D","C"], ["A","C","E"], ["B","C","D"], ["B","E","C","A"], ["D","E","C","A"]
    adjustment = {"B": [], "D": ["C"], "E": ["C", "A"]}
    _all = ["A", "B", "C", "D", "E", "F"]
    cost_table = {
        1: ({v: 1 for v in _all}, False),
        2: ({"A": 1, "B": 10, "C": 2, "D": 5, "E": 20, "F": 3}, False),
        3: ({"A": 1, "B": 10, "C": 2, "D": 5, "E": 20, "F": 3}, True),
        4: ({v: 1 for v in _all}, True),
    }

    def __init__(self, measurements, true_measurements=None):
        super().__init__(["B", "D", "E"])
        self.var_names = ["A", "B", "C", "D", "E", "F", "Y"]
        self.measurements = {v: np.asarray(measurements[v], np.float64).reshape(-1, 1) for v in self.var_names}

    def prior_columns(self, intervention_set):
        # keep the reference's column order for the two-variable sets that share C and A
        order = {("B",): ["B"], ("D",): ["D", "C"], ("E",): ["A", "C", "E"], ("B", "D"): ["B", "C", "D"],
                 ("B", "E"): ["B", "E", "C", "A"], ("D", "E"): ["D", "E", "C", "A"]}
        return list(order.get(tuple(intervention_set), super().prior_columns(intervention_set)))

    def fit_parameters_for(self, cols):
        return [1.0, 1.0, 10.0 if len(cols) >= 3 else 1.0, False]

    def define_sem(self):
        sem = OrderedDict()
        sem["U1"] = lambda epsilon, **kw: epsilon[0]
        sem["U2"] = lambda epsilon, **kw: epsilon[1]
        sem["F"] = lambda epsilon, **kw: epsilon[8]
        sem["A"] = lambda epsilon, U1, F, **kw: F ** 2 + U1 + epsilon[2]
        sem["B"] = lambda epsilon, U2, **kw: U2 + epsilon[3]
        sem["C"] = lambda epsilon, B, **kw: np.exp(-B) + epsilon[4]
        sem["D"] = lambda epsilon, C, **kw: np.exp(-C) / 10.0 + epsilon[5]
        sem["E"] = lambda epsilon, A, C, **kw: np.cos(A) + C / 10.0 + epsilon[6]
        sem["Y"] = lambda epsilon, D, E, U1, U2, **kw: (np.cos(D) - D / 5.0 + np.sin(E) - E / 4.0 + U1 + np.exp(-U2)
                                                         + epsilon[7])
        return sem

    def device_sem(self):
        """define_sem() as a program for cbo_sem_eval (term = (source, func, coef, scale): coef * func(scale * source))."""
        e = [f"e{i}" for i in range(9)]
        one = lambda src: (src, "id", 1.0, 1.0)
        return {"noise": e, "draw": self._draw_gaussian_noise(9),
                "nodes": [("U1", 0.0, [one("e0")]), ("U2", 0.0, [one("e1")]), ("F", 0.0, [one("e8")]),
                          ("A", 0.0, [("F", "square", 1.0, 1.0), one("U1"), one("e2")]),
                          ("B", 0.0, [one("U2"), one("e3")]),
                          ("C", 0.0, [("B", "exp", 1.0, -1.0), one("e4")]),
                          ("D", 0.0, [("C", "exp", 1.0 / 10.0, -1.0), one("e5")]),
                          ("E", 0.0, [("A", "cos", 1.0, 1.0), ("C", "id", 1.0 / 10.0, 1.0), one("e6")]),
                          ("Y", 0.0, [("D", "cos", 1.0, 1.0), ("D", "id", -1.0 / 5.0, 1.0), ("E", "sin", 1.0, 1.0),
                                      ("E", "id", -1.0 / 4.0, 1.0), one("U1"), ("U2", "exp", 1.0, -1.0), one("e7")])]}

    @staticmethod
    def get_exploration_set(set_name):
        mis = [["B"], ["D"], ["E"], ["B", "D"], ["B", "E"], ["D", "E"]]
        pomis = [["B"], ["D"], ["E"], ["B", "D"], ["D", "E"]]
        return mis if set_name == "MIS" else pomis

    @staticmethod
    def get_interventional_ranges():
        return OrderedDict([("E", [-6, 3]), ("B", [-5, 4]), ("D", [-5, 5]), ("F", [-4, 4])])
