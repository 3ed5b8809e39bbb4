"""Policy / data-preparation helpers of the agent (reference: src/utils_functions/cbo_functions.py)."""
import itertools

import numpy as np
import scipy.spatial


def _stack(samples, variables):
    return np.column_stack([np.asarray(samples[v], np.float64).reshape(-1) for v in variables])


def _hull(points):
    """Convex hull; degenerate (coplanar) inputs -- the standardised coral data has two perfectly correlated columns,
    on which the reference's plain Qhull call raises -- are joggled into general position."""
    try:
        return scipy.spatial.ConvexHull(points)
    except scipy.spatial.QhullError:
        return scipy.spatial.ConvexHull(points, qhull_options="QJ")


def update_hull(observational_samples, manipulative_variables):
    """Volume of the convex hull of the observed manipulative variables (reference :7-17)."""
    return _hull(_stack(observational_samples, manipulative_variables)).volume


def observe(num_observation, complete_dataset=None, initial_num_obs_samples=None):
    """The slice of new observations (reference :20-22; it returns the same rows on every call, Appendix B #8)."""
    return complete_dataset[initial_num_obs_samples:(initial_num_obs_samples + num_observation)]


def compute_coverage(observational_samples, manipulative_variables, dict_ranges):
    """(coverage ratio, hull of the observations, volume of the interventional box) (reference :25-41)."""
    box = list(itertools.product(*[dict_ranges[v] for v in manipulative_variables]))
    coverage_total = _hull(np.asarray(box, np.float64)).volume
    hull_obs = _hull(_stack(observational_samples, manipulative_variables))
    return hull_obs.volume / coverage_total, hull_obs, coverage_total


def define_initial_data_cbo(interventional_data, num_interventions, exploration_set, name_index, task):
    """Initial interventional data per exploration set: the shipped design of set j, shuffled with seed `name_index`
    (the global RNG state is saved and restored) and cut to `num_interventions` rows; plus the incumbent
    (reference :44-115).  interventional_data[j] = [k, name_1..name_k, X (p,k), y (p,1)]."""
    pick = np.min if task == "min" else np.max
    data_x_list, data_y_list, opt_list = [], [], []
    # Rows are matched to exploration sets by their variable NAMES: the reference matches by position
    # (cbo_functions.py:51-52), which hands a set another set's design whenever the shipped row order differs from the
    # exploration set in use (complete_graph MIS: SURVEY.md Appendix B #10; any POMIS list that skips a MIS entry).
    by_name = {}
    for row in interventional_data:
        k = int(row[0])
        by_name[tuple(str(v) for v in row[1:1 + k])] = row
    for j in range(len(exploration_set)):
        key = tuple(str(v) for v in exploration_set[j])
        perm = None
        if key in by_name:
            row = by_name[key]
        else:   # same variables in another order: take that row and permute its columns
            match = [kk for kk in by_name if sorted(kk) == sorted(key)]
            if not match:
                raise KeyError(f"no interventional data for exploration set {list(key)} (rows: {sorted(by_name)})")
            row = by_name[match[0]]
            perm = [match[0].index(v) for v in key]
        k = int(row[0])
        x = np.asarray(row[k + 1], np.float64)
        if perm is not None:
            x = x.reshape(len(x), -1)[:, perm]
        y = np.asarray(row[-1], np.float64)
        x = x.reshape(len(x), -1)
        y = y.reshape(len(y), -1)[:, :1]
        both = np.concatenate((x, y), axis=1)
        state = np.random.get_state()
        np.random.seed(name_index)
        np.random.shuffle(both)
        np.random.set_state(state)
        both = both[:num_interventions]
        data_x_list.append(both[:, :-1])
        data_y_list.append(both[:, -1:])
        opt_list.append(pick(both[:, -1]))
    opt_y = pick(opt_list)
    j_best = int(np.where(np.asarray(opt_list) == opt_y)[0][0])
    best_variable = "".join(exploration_set[j_best])
    xb, yb = data_x_list[j_best], data_y_list[j_best][:, 0]
    best_intervention_value = xb[yb == pick(yb)][0]
    return data_x_list, data_y_list, best_intervention_value, opt_y, best_variable
