"""Stand-in for the slice of GPy 1.10 that ChampiB/CBO_with_OOP imports.  See ../README.md."""
from . import core, kern, models, util  # noqa: F401
