class GridRBF:
    """Imported by causal_kernels.py:6; only used by get_one_dimensional_kernel, which the path never calls."""

    def __init__(self, *a, **kw):
        raise NotImplementedError("GridRBF is outside the acquisition path")
