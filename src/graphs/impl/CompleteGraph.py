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

    @staticmethod
    def get_exploration_set(set_name):
        mis = [["B"], ["D"], ["E"], ["B", "D"], ["B", "E"], ["D", "E"]]
        pomis = [["B"], ["D"], ["E"], ["B", "D"], ["D", "E"]]
        return mis if set_name == "MIS" else pomis

    @staticmethod
    def get_interventional_ranges():
        return OrderedDict([("E", [-6, 3]), ("B", [-5, 4]), ("D", [-5, 5]), ("F", [-4, 4])])
