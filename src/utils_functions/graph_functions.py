"""Structural-equation sampling and the parameter space (reference: src/utils_functions/graph_functions.py).

compute_interventions is the ground-truth simulator of a run, not part of the acquisition sweep; it is evaluated
on arrays (all Monte-Carlo samples in one pass) instead of the reference's 100 000-iteration Python loop.  For SEMs
whose noise comes only from `epsilon` (toy, complete graph) the draws are the same stream as the reference's."""
import copy

import numpy as np
from numpy.random import randn


class ContinuousParameter:
    """Minimal stand-in for emukit.core.ContinuousParameter."""

    def __init__(self, name, min_value, max_value):
        self.name, self.min, self.max = name, float(min_value), float(max_value)

    @property
    def bounds(self):
        return [(self.min, self.max)]


class ParameterSpace:
    """Minimal stand-in for emukit.core.ParameterSpace: an ordered list of continuous parameters."""

    def __init__(self, parameters):
        self.parameters = list(parameters)
        self.constraints = []

    @property
    def parameter_names(self):
        return [p.name for p in self.parameters]

    def get_bounds(self):
        return [(p.min, p.max) for p in self.parameters]

    def sample_uniform(self, point_count):
        b = np.asarray(self.get_bounds())
        return np.random.uniform(b[:, 0], b[:, 1], size=(point_count, len(b)))

    def grid_tables(self, points_per_dim=100):
        """Per-dimension np.linspace coordinate tables of the candidate grid."""
        return [np.linspace(p.min, p.max, points_per_dim) for p in self.parameters]


def sample_from_model(model, epsilon=None):
    """One sample (or, for a 2-D epsilon of shape (nodes, n), n samples at once) from a SEM (reference :8-27)."""
    epsilon = randn(len(model)) if epsilon is None else epsilon
    sample = {}
    for variable, function in model.items():
        sample[variable] = function(epsilon, **sample)
    return sample


def intervene_dict(model, **interventions):
    """SEM with the intervened variables replaced by constants (reference :30-45)."""
    new_model = copy.copy(model)
    for k, v in interventions.items():
        new_model[k] = (lambda value: (lambda *args, **kwargs: value))(v)
    return new_model


def compute_interventions(model, interventions, node_values, target_variable="Y", num_samples=100000, seed=1):
    """E[target | do(interventions = node_values[0])] by Monte Carlo, shape (1, 1) (reference :48-77)."""
    for i, node in enumerate(interventions.keys()):
        interventions[node] = node_values[0, i]
    mutilated = intervene_dict(model, **interventions)
    np.random.seed(seed)
    eps = randn(num_samples, len(model)).T          # row i of the transpose = epsilon[i] for every sample
    samples = sample_from_model(mutilated, eps)
    target = np.broadcast_to(np.asarray(samples[target_variable], np.float64), (num_samples,))
    return np.asarray(np.mean(target))[np.newaxis, np.newaxis]


def get_parameter_space(interventions, min_interventions, max_interventions):
    """ParameterSpace of the intervened variables (reference :80-93)."""
    return ParameterSpace([ContinuousParameter(*p) for p in zip(interventions.keys(), min_interventions, max_interventions)])
