from .GraphInterface import GraphInterface
from src.graphs.impl.ToyGraph import ToyGraph
from src.graphs.impl.CompleteGraph import CompleteGraph
from src.graphs.impl.CoralGraph import CoralGraph
from src.graphs.impl.SimplifiedCoralGraph import SimplifiedCoralGraph
