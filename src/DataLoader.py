"""Measurements and causal graph of an experiment (reference: src/DataLoader.py).

Looks for the reference layout first (./data/<experiment>/observations.pkl, true_observations.pkl,
interventional_data.npy); without it, reads the pickle-free fixture tests/golden/data/<experiment>.npz that
tests/golden/import_reference_data.py derived from the shipped data.  Adds the toy_graph entry the reference lacks
and lets simplified_coral_graph borrow coral_graph's field measurements."""
import os

import numpy as np
import pandas as pd

from src.graphs import *  # noqa: F401,F403

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fixture(experiment):
    for root in (os.path.join(_REPO, "tests", "golden", "data"), os.path.join("tests", "golden", "data")):
        path = os.path.join(root, experiment + ".npz")
        if os.path.exists(path):
            return np.load(path, allow_pickle=False)
    return None


class DataLoader:
    graph_classes = {
        "toy_graph": ToyGraph,
        "complete_graph": CompleteGraph,
        "coral_graph": CoralGraph,
        "simplified_coral_graph": SimplifiedCoralGraph,
    }

    def __init__(self, experiment, initial_num_obs_samples):
        folder = os.path.join(".", "data", experiment)
        true_measurements = None
        if os.path.exists(os.path.join(folder, "observations.pkl")):
            self.all_measurements = pd.read_pickle(os.path.join(folder, "observations.pkl"))
            tpath = os.path.join(folder, "true_observations.pkl")
            if os.path.exists(tpath):
                true_measurements = pd.read_pickle(tpath)
            self.interventions = np.load(os.path.join(folder, "interventional_data.npy"), allow_pickle=True)
        else:
            z = _fixture(experiment)
            if z is None:
                raise FileNotFoundError(f"no data for experiment {experiment!r}: neither {folder} nor the npz fixture exists")
            self.all_measurements = pd.DataFrame(z["observations"], columns=[str(c) for c in z["columns"]])
            if "true_observations" in z:
                true_measurements = pd.DataFrame(z["true_observations"], columns=[str(c) for c in z["true_columns"]])
            rows = []
            for j in range(int(z["num_sets"])):
                names = [str(v) for v in z[f"set{j}_names"]]
                rows.append([len(names), *names, z[f"set{j}_x"], z[f"set{j}_y"]])
            self.interventions = np.empty(len(rows), dtype=object)
            for j, r in enumerate(rows):
                self.interventions[j] = r
        if true_measurements is None and experiment == "simplified_coral_graph":
            z = _fixture("coral_graph")
            coral = os.path.join(".", "data", "coral_graph", "true_observations.pkl")
            if os.path.exists(coral):
                true_measurements = pd.read_pickle(coral)
            elif z is not None:
                true_measurements = pd.DataFrame(z["true_observations"], columns=[str(c) for c in z["true_columns"]])
        self.measurements = self.all_measurements[:initial_num_obs_samples]
        arguments = [self.measurements] + ([true_measurements] if true_measurements is not None else [])
        self.graph = self.graph_classes[experiment](*arguments)
        self._align_interventions()

    def _align_interventions(self):
        """The agent matches interventional rows to exploration sets by POSITION (cbo_functions.py:51-52); the shipped
        complete_graph rows are ordered E, B, D, ... while MIS is B, D, E, ....  Reorder the
        rows by their variable names so that every set starts from its own data."""
        by_name = {}
        for row in self.interventions:
            k = int(row[0])
            by_name[tuple(str(v) for v in row[1:1 + k])] = row
        ordered = []
        for s in self.graph.get_exploration_set("MIS"):
            key = tuple(s)
            if key not in by_name:
                return                      # unknown layout: leave untouched
            ordered.append(by_name[key])
        out = np.empty(len(ordered), dtype=object)
        for j, r in enumerate(ordered):
            out[j] = r
        self.interventions = out
