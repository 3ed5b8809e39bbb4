"""CPU tests of the reference-shaped host code in `src/` (no device calls): data loading, the explicit set -> GP-column
table, initial interventional data, costs, incumbent search, SEM sampling, the agent's constructor and policy pieces."""
import os
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_args(**kw):
    base = dict(exploration_set="MIS", initial_num_obs_samples=100, num_interventions=10, type_cost=1,
                num_additional_observations=20, num_trials=4, name_index=0, seed=9, causal_prior=True, experiment="complete_graph",
                task="min", grid_points=20, device="cuda:0", num_sem_samples=500,
                observational_fit="host")
    base.update(kw)
    return types.SimpleNamespace(**base)


@pytest.fixture()
def in_tmp(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    return tmp_path


def test_runcbo_imports_resolve():
    """The three import lines of the reference's runCBO.py."""
    ns = {}
    exec("from src.CBO import *\nfrom src.ArgumentParser import ArgumentParser as ArgumentParser\n"
         "from src.DataLoader import DataLoader as DataLoader\n", ns)
    assert "CBO" in ns and "ArgumentParser" in ns and "DataLoader" in ns
    from src.GaussianProcessFactory import GaussianProcessType
    assert (GaussianProcessType.GRAPH_GP, GaussianProcessType.CAUSAL_GP, GaussianProcessType.NON_CAUSAL_GP) == (0, 1, 2)


def test_argument_parser_defaults_and_seed():
    from src.ArgumentParser import ArgumentParser
    a = ArgumentParser().parse(argv=[])
    assert (a.initial_num_obs_samples, a.num_interventions, a.type_cost, a.num_additional_observations, a.num_trials,
            a.name_index, a.seed, a.exploration_set, a.causal_prior, a.experiment, a.task) == \
        (100, 10, 1, 20, 40, 0, 9, "MIS", False, "complete_graph", "min")
    first = np.random.rand()
    ArgumentParser().parse(argv=["--seed", "9"])
    assert np.random.rand() == first                       # numpy.random.seed(args.seed)
    assert ArgumentParser().parse(argv=["--causal_prior", "False"]).causal_prior is True   # type=bool quirk kept


@pytest.mark.parametrize("experiment,n_sets", [("toy_graph", 2), ("complete_graph", 6), ("coral_graph", 25),
                                               ("simplified_coral_graph", 25)])
def test_data_loader_and_prior_table(in_tmp, experiment, n_sets):
    from src.DataLoader import DataLoader
    data = DataLoader(experiment, 100)
    g = data.graph
    sets = g.get_exploration_set("MIS")
    assert len(sets) == n_sets and len(data.interventions) == n_sets and data.measurements.shape[0] == 100
    for j, s in enumerate(sets):
        row = data.interventions[j]
        assert [str(v) for v in row[1:1 + int(row[0])]] == s            # rows aligned with the exploration sets
        cols = g.prior_columns(s)
        assert cols[:len(s)] == s or set(s) <= set(cols)
        assert len(set(cols)) == len(cols) <= 8 and "Y" not in cols
        assert all(c in data.measurements.columns for c in cols)
    if experiment == "complete_graph":      # the reference's fit_dependencies rows (CompleteGraph.py:31-42)
        assert [g.prior_columns(s) for s in sets] == [["B"], ["D", "C"], ["A", "C", "E"], ["B", "C", "D"],
                                                      ["B", "E", "C", "A"], ["D", "E", "C", "A"]]
    if experiment == "coral_graph":         # rows the reference has (CoralGraph.py:54-70) and ones it lacks
        assert g.prior_columns(["N", "O"]) == ["N", "O", "S", "T", "D", "TE"]
        assert g.prior_columns(["O", "C"]) == ["O", "C", "S", "T", "D", "TE", "N", "L"]
        assert g.prior_columns(["N", "C"]) == ["N", "C", "L", "TE"]
        assert g.get_gp_name(g.prior_columns(["T"])) == "gp_T_S"


def test_initial_interventional_data_is_a_seeded_subset(in_tmp):
    from src.DataLoader import DataLoader
    from src.utils_functions import define_initial_data_cbo
    data = DataLoader("complete_graph", 100)
    sets = data.graph.get_exploration_set("MIS")
    state = np.random.get_state()[1].copy()
    xs, ys, best_x, opt_y, best_var = define_initial_data_cbo(data.interventions, 10, sets, 0, "min")
    np.testing.assert_array_equal(np.random.get_state()[1], state)          # global RNG state restored
    xs2, ys2, *_ = define_initial_data_cbo(data.interventions, 10, sets, 0, "min")
    for j, s in enumerate(sets):
        assert xs[j].shape == (10, len(s)) and ys[j].shape == (10, 1)
        np.testing.assert_array_equal(xs[j], xs2[j])
        full = np.hstack([np.asarray(data.interventions[j][len(s) + 1]).reshape(20, -1), np.asarray(data.interventions[j][-1]).reshape(20, 1)])
        for row in np.hstack([xs[j], ys[j]]):
            assert np.any(np.all(np.isclose(full, row), axis=1))
    assert opt_y == min(y.min() for y in ys) and best_var in ["".join(s) for s in sets]
    xs3, *_ = define_initial_data_cbo(data.interventions, 10, sets, 1, "min")
    assert any(not np.array_equal(a, b) for a, b in zip(xs, xs3))


def test_costs_and_incumbent():
    from src.graphs import CompleteGraph
    from src.utils_functions import Cost, find_current_global, total_cost
    meas = {v: np.zeros(5) for v in ["A", "B", "C", "D", "E", "F", "Y"]}
    g = CompleteGraph(meas)
    c1, c3 = g.get_cost_structure(1), g.get_cost_structure(3)
    X = np.array([[1.0, -2.0], [0.5, 0.5]])
    assert Cost(c1, ["B", "E"]).evaluate(X) == 2
    assert Cost(c3, ["B", "E"]).evaluate(X) == 10 + 1.5 + 20 + 2.5          # the reference sums |x| over the batch (#12)
    assert Cost(c1, ["B", "E"]).kernel_form() == (2.0, False) and Cost(c3, ["B", "E"]).kernel_form() == (30.0, True)
    assert g.fixed_cost_of(["B", "E"], 2) == (30.0, False)
    assert total_cost(["B", "D"], c3, {"B": -1.0, "D": 2.0}) == 10 + 1 + 5 + 2
    with pytest.raises(RuntimeError):
        g.get_cost_structure(7)
    cur = {"B": [np.inf, 0.4], "D": [np.inf], "BD": [np.inf, -0.2, 0.1]}
    assert find_current_global(cur, ["B", "D", "BD"], "min") == -0.2
    cur = {k: [-np.inf if x == np.inf else x for x in v] for k, v in cur.items()}
    assert find_current_global(cur, ["B", "D", "BD"], "max") == 0.4


def test_sem_sampling_matches_shipped_causal_effects(in_tmp):
    """compute_interventions (vectorised Monte Carlo over the SEM) against the true effects shipped with the reference."""
    from src.DataLoader import DataLoader
    from src.utils_functions import compute_interventions
    data = DataLoader("toy_graph", 100)
    sem = data.graph.define_sem()
    xz, yz = data.interventions[1][2], data.interventions[1][3]
    for i in (3, 10, 17):
        y = compute_interventions(sem, {"Z": ""}, xz[i:i + 1], num_samples=200000)
        assert y.shape == (1, 1) and abs(y[0, 0] - yz[i, 0]) < 0.02
    data = DataLoader("complete_graph", 100)
    sem = data.graph.define_sem()
    j = 3                                                                    # set ['B', 'D']
    xb, yb = data.interventions[j][3], data.interventions[j][4]
    y = compute_interventions(sem, {"B": "", "D": ""}, xb[5:6], num_samples=200000)
    assert abs(y[0, 0] - yb[5, 0]) < 0.05


def test_agent_constructor_and_policy_pieces(in_tmp):
    from src.CBO import CBO
    from src.DataLoader import DataLoader
    args = make_args()
    np.random.seed(args.seed)
    cbo = CBO(args, DataLoader("complete_graph", 100), verbose=False)
    assert cbo.es_size == 6 and cbo.intervention_names == ["B", "D", "E", "BD", "BE", "DE"] and cbo.max_n == 150
    assert cbo.saving_dir == "./data/complete_graph/fix_equal/100/10/" and os.path.isdir(cbo.saving_dir)
    assert 0 < cbo.epsilon < 10
    assert cbo.get_new_observation().shape[0] == 20
    assert len(cbo.monitor.space_list) == 6 and cbo.monitor.space_list[3].parameter_names == ["B", "D"]
    assert cbo.monitor.space_list[3].get_bounds() == [(-5.0, 4.0), (-5.0, 5.0)]
    tabs = cbo.monitor.space_list[3].grid_tables(20)
    np.testing.assert_array_equal(tabs[0], np.linspace(-5, 4, 20))
    ys = [np.array([[0.1]]), np.array([[0.7]]), np.array([[0.7]]), np.array([[np.nan]]), np.array([[0.2]]), np.array([[0.0]])]
    s, i = cbo.select_next_intervention(ys)
    assert (s, i) == (["D"], 1) and cbo.monitor.last_intervention == 1      # first maximum, NaN never wins
    xs = [np.zeros((1, len(v))) + 2.0 for v in cbo.exploration_set]
    assert cbo.compute_cost(["B", "D"], 3, xs) == 2.0
    # observe(): hyper-parameter fit on the host, closures only -- no device needed until they are called
    cbo.observe()
    assert len(cbo.mean_functions) == 6 and callable(cbo.mean_functions[0]) and cbo.measurements.shape[0] == 120
    assert cbo.monitor.type_trial == [0] and cbo.monitor.global_opt[-1] == cbo.monitor.global_opt[0]
    pr = cbo.do_calculus.set_problem(4)                                      # ['B', 'E'] -> GP on B, E, C, A
    assert pr.x_obs_int.shape == (120, 2) and pr.x_obs_cond.shape == (120, 2) and pr.g_total == 400 and pr.cost_fix == 2.0


def test_interventional_rows_follow_the_exploration_set_by_name(in_tmp):
    """The initial interventional data of every exploration set is the shipped row with the same variable names, whatever
    list of sets is in use (reference matches by position, cbo_functions.py:51-52: with complete_graph's POMIS list the
    set ['D', 'E'] would start from the ['B', 'E'] design)."""
    from src.CBO import CBO
    from src.DataLoader import DataLoader
    data = DataLoader("complete_graph", 100)
    shipped = {tuple(str(v) for v in row[1:1 + int(row[0])]): np.asarray(row[int(row[0]) + 1], np.float64).reshape(len(row[-1]), -1)
               for row in data.interventions}
    for es in ("MIS", "POMIS"):
        np.random.seed(9)
        cbo = CBO(make_args(exploration_set=es), data, verbose=False)
        assert cbo.exploration_set == data.graph.get_exploration_set(es)
        for s, variables in enumerate(cbo.exploration_set):
            design = shipped[tuple(variables)]
            x = np.asarray(cbo.monitor.data_x[s])
            assert x.shape == (10, len(variables))
            for row in x:       # every initial row of the set comes from the set's OWN shipped design
                assert np.any(np.all(np.isclose(design, row[None, :]), axis=1)), (es, variables)
    from src.utils_functions.cbo_functions import define_initial_data_cbo
    with pytest.raises(KeyError):
        define_initial_data_cbo(data.interventions, 10, [["B"], ["F"]], 0, "min")
