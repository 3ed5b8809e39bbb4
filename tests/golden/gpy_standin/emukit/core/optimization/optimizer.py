class OptLbfgs:
    def __init__(self, *a, **kw):
        raise NotImplementedError("outside the acquisition path")


class OptTrustRegionConstrained(OptLbfgs):
    pass


def apply_optimizer(*a, **kw):
    raise NotImplementedError("outside the acquisition path")
