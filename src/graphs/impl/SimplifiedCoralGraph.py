"""Coral-reef graph on the standardised observational data (reference: src/graphs/impl/SimplifiedCoralGraph.py): same
structure as CoralGraph, different interventional ranges.  The reference cannot load this experiment because
simplified_coral_graph/true_observations.pkl is not shipped (QUESTIONS.md:2); DataLoader falls back to the coral
field measurements, the obvious stand-in."""
from collections import OrderedDict

from src.graphs.impl.CoralGraph import CoralGraph


class SimplifiedCoralGraph(CoralGraph):
    ranges = OrderedDict([("N", [-2, 5]), ("O", [3, 4]), ("C", [0.3, 0.4]), ("T", [2300, 2400]), ("D", [2000, 2080])])
