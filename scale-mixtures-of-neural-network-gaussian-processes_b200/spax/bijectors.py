"""Positive-constraint bijectors, mirroring spax/bijectors.py:21-62 (host scalars, NumPy)."""
import numpy as np

__all__ = ["positive", "Softplus", "Exp"]


class PositiveBijector:
    def __init__(self, lower: float = 0.):
        self.lower = lower

    def __call__(self, x):
        return self.lower + self.base(x)

    def inverse(self, x):
        return self.base_inv(x - self.lower)

    def grad(self, x):
        """d (constrained value) / d (unconstrained value) - the chain-rule factor reverse-mode AD applies when the
        reference differentiates through ConstraintTrainVar.safe_value (spax/base.py:23-25)."""
        return self.base_grad(x)


class Exp(PositiveBijector):
    base = staticmethod(lambda x: np.exp(x))
    base_inv = staticmethod(lambda x: np.log(x))
    base_grad = staticmethod(lambda x: np.exp(x))


class Softplus(PositiveBijector):
    base = staticmethod(lambda x: np.logaddexp(x, 0.0))                                  # jax.nn.softplus
    base_inv = staticmethod(lambda x: np.where(x < 20., np.log(np.expm1(np.minimum(x, 20.))), x))
    base_grad = staticmethod(lambda x: 1.0 / (1.0 + np.exp(-x)))                           # sigmoid


def positive(lower=None, base=None):
    lower_bound = lower if lower is not None else 0.0
    name = (base if base is not None else "softplus").lower()
    return {"exp": Exp, "softplus": Softplus}[name](lower_bound)
