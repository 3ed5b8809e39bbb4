class AcquisitionOptimizerBase:
    def __init__(self, *a, **kw):
        raise NotImplementedError("AcquisitionOptimizerBase: the anchor-point / L-BFGS optimiser is outside the acquisition path")
