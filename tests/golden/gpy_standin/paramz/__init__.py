"""Stand-in for the slice of paramz 0.9.5 the reference touches (Param with .fix(value), Logexp)."""
import numpy as np

from . import transformations  # noqa: F401


class Param(np.ndarray):
    """A named float64 array.  ``fix(value)`` is paramz's ``constrain_fixed(value=None, ...)``: when a value is given the
    parameter is overwritten with it (this is what utils.py:43 ``gp.likelihood.variance.fix(1e-2)`` relies on)."""

    def __new__(cls, name, input_array, default_constraint=None, *a, **kw):
        obj = np.atleast_1d(np.array(input_array, dtype=np.float64)).view(cls)
        obj.name = name
        obj.constraint = default_constraint
        obj.is_fixed = False
        obj.gradient = np.zeros(obj.shape)
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.name = getattr(obj, "name", None)
        self.constraint = getattr(obj, "constraint", None)
        self.is_fixed = getattr(obj, "is_fixed", False)
        self.gradient = getattr(obj, "gradient", None)

    def constrain_fixed(self, value=None, warning=True, trigger_parent=True):
        if value is not None:
            self[:] = value
        self.is_fixed = True
        return self

    fix = constrain_fixed

    def copy(self):
        return Param(self.name, np.array(self), self.constraint)


class Parameterized:
    def __init__(self, name=None, *a, **kw):
        self.name = name
        self.parameters = []

    def link_parameter(self, p, index=None):
        self.parameters.append(p)

    def link_parameters(self, *ps):
        for p in ps:
            self.link_parameter(p)

    def unlink_parameter(self, p):
        self.parameters = [q for q in self.parameters if q is not p]

    def parameters_changed(self):
        pass
