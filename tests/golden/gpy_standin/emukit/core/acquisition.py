"""emukit.core.acquisition.Acquisition and its operator overloads (utils.py:34 divides two acquisitions)."""


class Acquisition:
    def evaluate(self, x):
        raise NotImplementedError

    @property
    def has_gradients(self):
        raise NotImplementedError

    def evaluate_with_gradients(self, x):
        raise NotImplementedError

    def __add__(self, other):
        return Sum(self, other)

    def __mul__(self, other):
        return Product(self, other)

    def __rmul__(self, other):
        return Product(other, self)

    def __truediv__(self, denominator):
        return Quotient(self, denominator)


class Quotient(Acquisition):
    def __init__(self, numerator, denominator):
        self.numerator, self.denominator = numerator, denominator

    def evaluate(self, x):
        return self.numerator.evaluate(x) / self.denominator.evaluate(x)

    @property
    def has_gradients(self):
        return self.numerator.has_gradients and self.denominator.has_gradients


class Product(Acquisition):
    def __init__(self, a, b):
        self.a, self.b = a, b

    def evaluate(self, x):
        return self.a.evaluate(x) * self.b.evaluate(x)


class Sum(Acquisition):
    def __init__(self, a, b):
        self.a, self.b = a, b

    def evaluate(self, x):
        return self.a.evaluate(x) + self.b.evaluate(x)
