"""Bookkeeping of a CBO run -- the role of the reference's src/Monitor.py, kept API-compatible because the agent
(src/CBO.py), the tests and downstream analysis scripts read its attributes and files by name.

Design: the run is a list of `Trial` records (one per observe / intervene call, plus the initial state) and every list
the reference exposes (`type_trial`, `current_cost`, `global_opt`, `observed`, `trial_intervened`, `cumulative_cost`) is
DERIVED from that list by a property, so the log cannot get out of step with itself.  The per-set state (`data_x`,
`data_y`, `current_best_x`, `current_best_y`, `space_list`, `target_function_list`) lives in `SetState` objects and is
exposed under the reference's names as read-through lists / dicts.  Ground truth of an intervention: the graph's SEM
on the GPU (cbo_with_oop_b200/sem.py) when the graph provides a device program and a CUDA device is
present, else the vectorised host Monte Carlo (graph_functions.compute_interventions) -- both reseed with seed 1 per
call like the reference (graph_functions.py:73).
Out of the accelerated path; touches it only at add_intervention_data (reference :148-160)."""
from __future__ import annotations

import copy
import time
from dataclasses import dataclass, field
from functools import partial
from typing import Dict, List, Optional

import numpy as np

from src.utils_functions import *  # noqa: F401,F403


@dataclass
class Trial:
    """One step of the agent.  kind: 'init' | 'observe' | 'intervene'."""
    kind: str
    incumbent: float                 # best target value known after the step
    cost: float                      # cumulative intervention cost after the step
    set_index: Optional[int] = None  # exploration set intervened on
    x: Optional[np.ndarray] = None   # intervention values
    y: Optional[float] = None        # target value the SEM returned


@dataclass
class SetState:
    """Everything the run keeps per exploration set."""
    name: str
    variables: List[str]
    x: np.ndarray                    # interventional inputs so far (n, d)
    y: np.ndarray                    # interventional outputs so far (n, 1)
    space: object
    target: object                   # callable x (1, d) -> y (1, 1): the ground-truth simulator
    best_x: list = field(default_factory=list)
    best_y: list = field(default_factory=list)


class _SetListView:
    """List-like view of one attribute of the SetState objects (monitor.data_x[s] = ..., len(), iteration)."""

    def __init__(self, sets, attr):
        self._sets, self._attr = sets, attr

    def __getitem__(self, i):
        return getattr(self._sets[i], self._attr)

    def __setitem__(self, i, value):
        setattr(self._sets[i], self._attr, value)

    def __len__(self):
        return len(self._sets)

    def __iter__(self):
        return (getattr(s, self._attr) for s in self._sets)


class Monitor:
    def __init__(self, cbo, verbose=False):
        self.cbo, self.verbose = cbo, verbose
        xs, ys, best_value, opt_y, best_variable = define_initial_data_cbo(
            cbo.interventions, cbo.num_interventions, cbo.exploration_set, cbo.name_index, cbo.task)
        worst = np.inf if cbo.task == "min" else -np.inf
        ranges, sem = cbo.graph.get_interventional_ranges(), cbo.graph.define_sem()
        self.sets: List[SetState] = []
        for s, variables in enumerate(cbo.exploration_set):
            slots = {v: "" for v in variables}
            space = get_parameter_space(slots, [ranges[v][0] for v in variables], [ranges[v][1] for v in variables])
            target = self._make_target(sem, slots, variables)
            name = cbo.intervention_names[s]
            st = SetState(name, list(variables), xs[s], ys[s], space, target, best_x=[worst], best_y=[worst])
            if name == best_variable:
                st.best_x.append(best_value)
                st.best_y.append(opt_y)
            self.sets.append(st)
        self.trials: List[Trial] = [Trial("init", opt_y, 0.0)]
        self.last_intervention = None
        self.start_time = self.total_time = None

    # ---- ground truth ----------------------------------------------------------------------------------
    def _make_target(self, sem, slots, variables):
        """E[Y | do(variables = x)] of the graph's SEM.  Device path: one kernel over all Monte-Carlo samples."""
        cbo = self.cbo
        host = partial(compute_interventions, sem, slots, target_variable="Y", num_samples=cbo.num_sem_samples)
        program = getattr(cbo.graph, "device_sem", None)
        if program is None or getattr(cbo, "ground_truth", "device") != "device":
            return host
        state = {"sim": None}

        def target(x):
            import torch
            if not torch.cuda.is_available():
                return host(x)
            if state["sim"] is None:
                from cbo_with_oop_b200.sem import DeviceSEM
                state["sim"] = DeviceSEM(cbo.graph, num_samples=cbo.num_sem_samples, device=cbo.device)
            return state["sim"].mean_target(list(variables), np.asarray(x, np.float64).reshape(1, -1)).reshape(1, 1)
        return target

    # ---- the reference's attribute names, derived from the records ---------------------------------------
    @property
    def type_trial(self):
        return [1 if t.kind == "intervene" else 0 for t in self.trials[1:]]

    @property
    def i(self):
        return len(self.trials) - 1

    @property
    def observed(self):
        return sum(t.kind == "observe" for t in self.trials)

    @property
    def trial_intervened(self):
        return float(sum(t.kind == "intervene" for t in self.trials))

    @property
    def cumulative_cost(self):
        return self.trials[-1].cost

    @property
    def current_cost(self):
        return [t.cost for t in self.trials]

    @property
    def global_opt(self):
        return [t.incumbent for t in self.trials]

    @property
    def data_x(self):
        return _SetListView(self.sets, "x")

    @property
    def data_y(self):
        return _SetListView(self.sets, "y")

    @property
    def space_list(self):
        return [s.space for s in self.sets]

    @property
    def target_function_list(self):
        return [s.target for s in self.sets]

    @property
    def current_best_x(self) -> Dict[str, list]:
        return {s.name: s.best_x for s in self.sets}

    @property
    def current_best_y(self) -> Dict[str, list]:
        return {s.name: s.best_y for s in self.sets}

    # ---- timing ------------------------------------------------------------------------------------------
    def start(self):
        self.start_time = time.time()

    def stop(self):
        self.total_time = time.time() - self.start_time

    # ---- logging (reference :82-139) ------------------------------------------------------------------------
    def log_agent_behaviour(self, act):
        """Open the record of this trial; log_agent_performance completes it."""
        if self.verbose:
            print("Optimization step", self.i)
        last = self.trials[-1]
        self._open = Trial("intervene" if act else "observe", last.incumbent, last.cost)

    def agent_previously_observed(self):
        """Was the trial BEFORE the one in progress an observation?  (reference: type_trial[-2] with the current trial appended)"""
        done = self.type_trial
        return (done[-1] if done else 0) == 0

    def log_agent_performance(self, intervention_set=None, intervention=None, acquisition_xs=None, current_cost=None):
        rec = self._open
        if current_cost is not None:           # an intervention: evaluate the target, extend the set's data, update the incumbent
            x = acquisition_xs[intervention]
            y = self.compute_target_function(intervention_set, intervention, acquisition_xs)
            self.add_intervention_data(y, intervention, acquisition_xs)
            st = self.sets[intervention]
            st.best_x.append(x[0][0])
            st.best_y.append(y[0][0])
            rec.set_index, rec.x, rec.y = intervention, np.array(x, copy=True), float(y[0][0])
            rec.incumbent = find_current_global(self.current_best_y, self.cbo.intervention_names, self.cbo.task)
            rec.cost = self.cumulative_cost + current_cost
            if self.verbose:
                print("####### Current_global #########", rec.incumbent)
        self.trials.append(rec)

    def add_intervention_data(self, target_ys, intervention, acquisition_xs):
        """Append (x*, y_new) to the chosen set and hand the data to its model (reference :148-160)."""
        st = self.sets[intervention]
        st.x = np.vstack((st.x, acquisition_xs[intervention]))
        st.y = np.vstack((st.y, target_ys))
        self.cbo.models[intervention].set_data(st.x, st.y)

    def compute_target_function(self, intervention_set, intervention, acquisition_xs):
        y_new = self.sets[intervention].target(acquisition_xs[intervention])
        if self.verbose:
            print("Selected intervention set: ", intervention_set)
            print("Selected values: ", acquisition_xs[intervention])
            print("Target function at the selected values: ", y_new)
        return y_new

    # ---- results (file names and contents of the reference, Monitor.py:181-191) -----------------------------
    def save_results(self):
        tag = f"{self.cbo.exploration_set}_{self.cbo.gp_type}_{self.cbo.name_index}"
        payload = {"cost": self.current_cost, "best_x": copy.deepcopy(self.current_best_x),
                   "best_y": copy.deepcopy(self.current_best_y), "total_time": self.total_time, "observed": self.observed,
                   "global_opt": self.global_opt}
        for name, value in payload.items():
            np.save(f"{self.cbo.saving_dir}{name}_{tag}.npy", value)
