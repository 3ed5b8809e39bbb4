"""Common behaviour of the causal graphs (reference: src/graphs/GraphInterface.py).

New in this build: `prior_columns(intervention_set)` -- the explicit table "exploration set -> columns of the
observational GP whose Monte-Carlo average is the causal prior" that the reference leaves implicit and partly
broken.  A graph lists the adjustment columns of every manipulative variable
(`adjustment`); the GP of a set regresses Y on the set's variables followed by the union of their adjustment
columns.  For the sets that do have a `fit_dependencies` row in the reference this reproduces that row.
"""
from __future__ import annotations

import abc
from collections import OrderedDict
from functools import partial
from typing import Dict, List

import numpy as np


class GraphInterface(abc.ABC):
    #: adjustment (conditioning) columns per manipulative variable; subclasses fill it in
    adjustment: Dict[str, List[str]] = {}
    #: fixed cost per variable for cost types 1..4 -> (dict of fixed costs, variable flag)
    cost_table: Dict[int, tuple] = {}

    def __init__(self, manipulative_variables):
        self.manipulative_variables = list(manipulative_variables)

    # ---- to be provided by each graph ---------------------------------------------------------------
    @abc.abstractmethod
    def define_sem(self):
        """OrderedDict name -> f(epsilon, **parents) (reference: define_sem)."""

    @staticmethod
    @abc.abstractmethod
    def get_exploration_set(set_name):
        ...

    @staticmethod
    @abc.abstractmethod
    def get_interventional_ranges():
        ...

    # ---- the SEM as a device program (cbo_with_oop_b200/sem.py) -------------------------
    @staticmethod
    def _draw_gaussian_noise(num_nodes):
        """Noise of a SEM whose only randomness is `epsilon`: the host function's draw, np.random.seed(seed) followed by
        randn(num_samples, num_nodes) (graph_functions.compute_interventions), one row per node."""
        def draw(num_samples, intervened, seed):
            np.random.seed(seed)
            return np.ascontiguousarray(np.random.randn(num_samples, num_nodes).T)
        return draw

    # ---- naming helpers (reference GraphInterface.py:29-43) -----------------------------------------
    @staticmethod
    def get_function_name(interventions):
        return "compute_do_" + "".join(interventions)

    @staticmethod
    def get_gp_name(interventions):
        return "gp_" + "_".join(interventions)

    # ---- costs (reference GraphInterface.py:46-50 and each graph's get_cost_structure) ---------------
    @staticmethod
    def cost(fix_cost, variable_cost, intervention_value, **kwargs):
        total = fix_cost
        if variable_cost is True:
            total += np.sum(np.abs(intervention_value))
        return total

    def get_cost_structure(self, type_cost):
        if type_cost not in self.cost_table:
            raise RuntimeError(f"[ERROR] Invalid cost type: {type_cost}")
        fixed, variable = self.cost_table[type_cost]
        return OrderedDict((name, partial(self.cost, fix, variable)) for name, fix in fixed.items())

    def fixed_cost_of(self, intervention_set, type_cost):
        """(sum of the fixed costs of the set's variables, variable flag): the form the sweep kernel consumes."""
        fixed, variable = self.cost_table[type_cost]
        return float(sum(fixed[v] for v in intervention_set)), bool(variable)

    # ---- the explicit prior table --------------------------------------------------------------------
    def prior_columns(self, intervention_set) -> List[str]:
        cols = list(intervention_set)
        for v in intervention_set:
            for a in self.adjustment.get(v, []):
                if a not in cols:
                    cols.append(a)
        return cols

    def prior_gp_name(self, intervention_set) -> str:
        return self.get_gp_name(self.prior_columns(intervention_set))

    # ---- observational GPs ----------------------------------------------------------------------------
    def fit_all_gaussian_processes(self, measurements=None, device=None):
        """One observational GP (RBF, noise fixed to 1e-2, hyper-parameters optimised; reference utils.py:40-45)
        per exploration set of MIS, keyed by its gp name.  `measurements` may be a DataFrame or a dict of columns.
        `device` (the agent passes cbo.device): fit on the GPU -- see utils.fit_gaussian_process."""
        from src.utils_functions.utils import fit_gaussian_process
        data = self.measurements if measurements is None else {
            name: np.asarray(measurements[name], np.float64).reshape(-1, 1) for name in self.var_names}
        gps = {}
        for s in self.get_exploration_set("MIS"):
            cols = self.prior_columns(s)
            name = self.get_gp_name(cols)
            if name in gps:
                continue
            x = np.hstack([data[c] for c in cols])
            gps[name] = fit_gaussian_process(x, data["Y"], self.fit_parameters_for(cols), device=device)
            gps[name].columns = cols
        return gps

    def fit_parameters_for(self, cols):
        """[lengthscale, variance, noise (ignored: fixed to 1e-2), ARD] start values (reference fit_parameters)."""
        return [1.0, 1.0, 1.0, False]
