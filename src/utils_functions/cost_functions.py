"""Intervention costs (reference: src/utils_functions/cost_functions.py)."""
import numpy as np


class Cost:
    """Denominator of the acquisition: sum over the set's variables of their cost callables
    (reference cost_functions.py:11-17).  `kernel_form()` gives the (fixed sum, variable flag) pair the sweep
    kernel applies per candidate."""

    def __init__(self, costs_functions, evaluated_set):
        self.costs_functions = costs_functions
        self.evaluated_set = evaluated_set

    def evaluate(self, x):
        x = np.atleast_2d(np.asarray(x, np.float64))
        return sum(self.costs_functions[name](x[:, i]) for i, name in enumerate(self.evaluated_set))

    def kernel_form(self):
        fix, variable = 0.0, False
        for name in self.evaluated_set:
            f = self.costs_functions[name]
            at0, at1 = float(f(np.zeros(1))), float(f(np.ones(1)))
            fix += at0
            variable = variable or (at1 != at0)
        return fix, variable

    @property
    def has_gradients(self):
        return True

    def evaluate_with_gradients(self, x):
        return self.evaluate(x), np.zeros(np.shape(x))


def total_cost(intervention_variables, costs, x_new_dict):
    """Cost of one performed intervention (reference cost_functions.py:27-31)."""
    return float(sum(costs[v](x_new_dict[v]) for v in intervention_variables))
