"""Bookkeeping of a CBO run (reference: src/Monitor.py): initial interventional data, the per-set parameter spaces and
target functions, logs of cost / best value / trial type, saving of the results."""
import copy
import time
from functools import partial

import numpy as np

from src.utils_functions import *  # noqa: F401,F403


class Monitor:
    def __init__(self, cbo, verbose=False):
        self.cbo = cbo
        self.verbose = verbose
        # initial interventional data: seeded shuffle of the shipped designs (cbo_functions.define_initial_data_cbo)
        self.data_x, self.data_y, best_value, opt_y, best_variable = define_initial_data_cbo(
            cbo.interventions, cbo.num_interventions, cbo.exploration_set, cbo.name_index, cbo.task)
        self.current_cost = [0.0]
        self.global_opt = [opt_y]
        worst = np.inf if cbo.task == "min" else -np.inf
        self.current_best_x = {name: [worst] for name in cbo.intervention_names}
        self.current_best_y = copy.deepcopy(self.current_best_x)
        self.current_best_y[best_variable].append(opt_y)
        self.current_best_x[best_variable].append(best_value)

        self.observed = 0
        self.trial_intervened = 0.0
        self.cumulative_cost = 0.0
        self.target_function_list, self.space_list, self.type_trial = [], [], []
        ranges = cbo.graph.get_interventional_ranges()
        sem = cbo.graph.define_sem()
        for variables in cbo.exploration_set:
            lows = [ranges[v][0] for v in variables]
            highs = [ranges[v][1] for v in variables]
            interventions = {v: "" for v in variables}
            self.space_list.append(get_parameter_space(interventions, lows, highs))
            self.target_function_list.append(partial(compute_interventions, sem, interventions, target_variable="Y",
                                                     num_samples=cbo.num_sem_samples))
        self.i = 0
        self.last_intervention = None
        self.start_time = None
        self.total_time = None

    def start(self):
        self.start_time = time.time()

    def stop(self):
        self.total_time = time.time() - self.start_time

    def log_agent_behaviour(self, act):
        """Record whether this trial is an intervention (act=True) or an observation (reference :82-101)."""
        if self.verbose is True:
            print("Optimization step", self.i)
        if act is True:
            self.type_trial.append(1)
            self.trial_intervened += 1
        else:
            self.observed += 1
            self.type_trial.append(0)
        self.i += 1

    def log_agent_performance(self, intervention_set=None, intervention=None, acquisition_xs=None, current_cost=None):
        """After an observation the cost and incumbent repeat; after an intervention evaluate the target, extend the
        set's dataset and update cost and incumbent (reference :103-139)."""
        if current_cost is None:
            self.global_opt.append(self.global_opt[-1])
            self.current_cost.append(self.current_cost[-1])
            return
        target_ys = self.compute_target_function(intervention_set, intervention, acquisition_xs)
        self.add_intervention_data(target_ys, intervention, acquisition_xs)
        name = self.cbo.intervention_names[intervention]
        self.current_best_x[name].append(acquisition_xs[intervention][0][0])
        self.current_best_y[name].append(target_ys[0][0])
        current_best = find_current_global(self.current_best_y, self.cbo.intervention_names, self.cbo.task)
        self.global_opt.append(current_best)
        self.cumulative_cost += current_cost
        self.current_cost.append(self.cumulative_cost)
        if self.verbose is True:
            print("####### Current_global #########", current_best)

    def agent_previously_observed(self):
        return self.type_trial[-2] == 0

    def add_intervention_data(self, target_ys, intervention, acquisition_xs):
        """Append (x*, y_new) to the chosen set and hand the data to its model (reference :148-160)."""
        self.data_x[intervention] = np.vstack((self.data_x[intervention], acquisition_xs[intervention]))
        self.data_y[intervention] = np.vstack((self.data_y[intervention], target_ys))
        self.cbo.models[intervention].set_data(self.data_x[intervention], self.data_y[intervention])

    def compute_target_function(self, intervention_set, intervention, acquisition_xs):
        y_new = self.target_function_list[intervention](acquisition_xs[intervention])
        if self.verbose is True:
            print("Selected intervention set: ", intervention_set)
            print("Selected values: ", acquisition_xs[intervention])
            print("Target function at the selected values: ", y_new)
        return y_new

    def save_results(self):
        index = f"{self.cbo.exploration_set}_{self.cbo.gp_type}_{self.cbo.name_index}"
        d = self.cbo.saving_dir
        np.save(d + f"cost_{index}.npy", self.current_cost)
        np.save(d + f"best_x_{index}.npy", self.current_best_x)
        np.save(d + f"best_y_{index}.npy", self.current_best_y)
        np.save(d + f"total_time_{index}.npy", self.total_time)
        np.save(d + f"observed_{index}.npy", self.observed)
        np.save(d + f"global_opt_{index}.npy", self.global_opt)
