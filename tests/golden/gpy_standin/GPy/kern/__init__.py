from .src.stationary import Stationary  # noqa: F401
from .src.rbf import RBF  # noqa: F401
