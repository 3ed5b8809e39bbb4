class ObjectiveAnchorPointsGenerator:
    def __init__(self, *a, **kw):
        raise NotImplementedError("ObjectiveAnchorPointsGenerator: the anchor-point / L-BFGS optimiser is outside the acquisition path")
