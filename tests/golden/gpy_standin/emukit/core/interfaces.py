class IModel:
    pass


class IDifferentiable:
    pass
