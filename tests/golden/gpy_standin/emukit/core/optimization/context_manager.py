class ContextManager:
    def __init__(self, *a, **kw):
        raise NotImplementedError("ContextManager: the anchor-point / L-BFGS optimiser is outside the acquisition path")
