from . import acquisition, interfaces  # noqa: F401


class ContinuousParameter:
    def __init__(self, name, min_value, max_value):
        self.name, self.min, self.max = name, min_value, max_value

    @property
    def bounds(self):
        return [(self.min, self.max)]


class ParameterSpace:
    def __init__(self, parameters, constraints=None):
        self.parameters = list(parameters)
        self.constraints = constraints or []

    def get_bounds(self):
        return [b for p in self.parameters for b in p.bounds]

    @property
    def parameter_names(self):
        return [p.name for p in self.parameters]
