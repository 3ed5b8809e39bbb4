from .gp_regression import GPRegression  # noqa: F401
