"""Coral-reef graphs (reference: src/graphs/impl/CoralGraph.py and SimplifiedCoralGraph.py): eleven observed variables,
manipulative variables N, O, C, T, D, target Y; the SEM is a chain of linear regressions fitted on the field
measurements (`true_measurements`) with a gamma law for the light L and a 3-component Gaussian mixture for the
nutrients N.  The SEM functions are written on arrays so that compute_interventions evaluates all Monte-Carlo
samples in one pass (the reference loops 100 000 times in Python, graph_functions.py:48-77)."""
from collections import OrderedDict

import numpy as np

from src.graphs.GraphInterface import GraphInterface


def _count(epsilon):
    e = np.asarray(epsilon)
    return e.shape[-1] if e.ndim == 2 else 1


class CoralGraph(GraphInterface):
    # back-door adjustment columns of each manipulative variable (reference rows ["N"], ["O","S","T","D","TE"],
    # ["C","N","L","TE"], ["T","S"], ["D","S"]); unions over a set reproduce the reference's multi-variable rows and
    # define the ten sets the reference has no row for
    adjustment = {"N": [], "O": ["S", "T", "D", "TE"], "C": ["N", "L", "TE"], "T": ["S"], "D": ["S"]}
    parents = OrderedDict([
        ("TE", ["L"]), ("C", ["N", "L", "TE"]), ("S", ["TE"]), ("T", ["S"]), ("D", ["S"]),
        ("P", ["S", "T", "D", "TE"]), ("O", ["S", "T", "D", "TE"]), ("CO", ["S", "T", "D", "TE"]),
        ("Y", ["L", "N", "P", "O", "C", "CO", "TE"]),
    ])
    _manip = ["N", "O", "C", "T", "D"]
    cost_table = {
        1: ({v: 1 for v in _manip}, False),
        2: ({"N": 1, "O": 10, "C": 2, "T": 5, "D": 20}, False),
        3: ({"N": 1, "O": 10, "C": 2, "T": 5, "D": 20}, True),
        4: ({v: 1 for v in _manip}, True),
    }
    ranges = OrderedDict([("N", [-2, 5]), ("O", [2, 4]), ("C", [0, 1]), ("T", [2450, 2500]), ("D", [1950, 1965])])

    def __init__(self, measurements, true_measurements=None):
        super().__init__(self._manip)
        self.var_names = ["Y", "N", "CO", "T", "D", "P", "O", "S", "L", "TE", "C"]
        self.measurements = {v: np.asarray(measurements[v], np.float64).reshape(-1, 1) for v in self.var_names}
        true_measurements = measurements if true_measurements is None else true_measurements
        self.true_measurements = {v: np.asarray(true_measurements[v], np.float64).reshape(-1, 1) for v in self.var_names}
        self._fit_structural_equations()

    def _fit_structural_equations(self):
        from scipy.stats import gamma
        from sklearn.linear_model import LinearRegression
        from sklearn.mixture import GaussianMixture
        self.regressions = {}
        for child, pars in self.parents.items():
            x = np.hstack([self.true_measurements[p] for p in pars])
            self.regressions[child] = LinearRegression().fit(x, self.true_measurements[child])
        a, loc, scale = gamma.fit(self.true_measurements["L"])
        self.dist_Light = gamma(a=a, loc=loc, scale=scale)
        self.dist_nutrients_pc1 = GaussianMixture(n_components=3).fit(self.true_measurements["N"])

    def _linear(self, child):
        reg = self.regressions[child]
        coef, icpt = np.asarray(reg.coef_).reshape(-1), float(np.asarray(reg.intercept_).reshape(-1)[0])
        pars = self.parents[child]

        def f(epsilon, **kw):
            out = icpt
            for c, p in zip(coef, pars):
                out = out + c * kw[p]
            return out
        return f

    def define_sem(self):
        def f_n(epsilon, **kw):
            n = _count(epsilon)
            s = self.dist_nutrients_pc1.sample(n)[0][:, 0]
            return s if n > 1 else s[0]

        def f_l(epsilon, **kw):
            n = _count(epsilon)
            s = self.dist_Light.rvs(n)
            return s if n > 1 else s[0]

        sem = OrderedDict([("N", f_n), ("L", f_l)])
        for child in ["TE", "C", "S", "T", "D", "P", "O", "CO", "Y"]:
            sem[child] = self._linear(child)
        return sem

    def device_sem(self):
        """define_sem() as a program for cbo_sem_eval.  The two stochastic roots are per-sample inputs drawn on the host
        exactly as compute_interventions draws them: np.random.seed(seed), randn(num_samples, 11) (unused by the linear
        equations but it advances the stream), then N from the Gaussian mixture unless N is intervened, then L from the
        gamma law -- both through the global NumPy stream, as define_sem()'s f_n / f_l do."""
        def draw(num_samples, intervened, seed):
            np.random.seed(seed)
            np.random.randn(num_samples, 11)
            n = np.zeros(num_samples) if "N" in intervened else self.dist_nutrients_pc1.sample(num_samples)[0][:, 0]
            light = self.dist_Light.rvs(num_samples)
            return np.vstack([n, light])

        nodes = [("N", 0.0, [("n_draw", "id", 1.0, 1.0)]), ("L", 0.0, [("l_draw", "id", 1.0, 1.0)])]
        for child in ["TE", "C", "S", "T", "D", "P", "O", "CO", "Y"]:
            reg = self.regressions[child]
            coef, icpt = np.asarray(reg.coef_).reshape(-1), float(np.asarray(reg.intercept_).reshape(-1)[0])
            nodes.append((child, icpt, [(p, "id", float(c), 1.0) for c, p in zip(coef, self.parents[child])]))
        return {"noise": ["n_draw", "l_draw"], "draw": draw, "nodes": nodes}

    @staticmethod
    def get_exploration_set(set_name):
        mis_1 = [["N"], ["O"], ["C"], ["T"], ["D"]]
        mis_2 = [["N", "O"], ["N", "C"], ["N", "T"], ["N", "D"], ["O", "C"], ["O", "T"], ["O", "D"], ["T", "C"], ["T", "D"],
                 ["C", "D"]]
        mis_3 = [["N", "O", "C"], ["N", "O", "T"], ["N", "O", "D"], ["N", "C", "T"], ["N", "C", "D"], ["N", "T", "D"],
                 ["O", "C", "T"], ["O", "C", "D"], ["C", "T", "D"], ["O", "T", "D"]]
        return mis_1 + mis_2 + mis_3          # POMIS == MIS in the reference ("To change")

    @classmethod
    def get_interventional_ranges(cls):
        return OrderedDict((k, list(v)) for k, v in cls.ranges.items())

    def fit_parameters_for(self, cols):
        # reference fit_parameters: ARD for the single-variable GPs with adjustment columns, scalar lengthscale otherwise
        single = {("O", "S", "T", "D", "TE"), ("C", "N", "L", "TE"), ("T", "S"), ("D", "S")}
        return [1.0, 1.0, 1.0, tuple(cols) in single]
