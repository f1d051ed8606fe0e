"""Module / variable shells mirroring spax/base.py:15-28 without objax: a ConstraintTrainVar stores the
unconstrained value and exposes the constrained ``safe_value``."""
import numpy as np

__all__ = ["Module", "TrainVar", "ConstraintTrainVar"]


class TrainVar:
    def __init__(self, tensor):
        self._value = np.asarray(tensor, dtype=np.float64)

    @property
    def value(self):
        return self._value

    def assign(self, tensor):
        self._value = np.asarray(tensor, dtype=np.float64)


class ConstraintTrainVar(TrainVar):
    def __init__(self, tensor, constraint):
        super().__init__(constraint.inverse(np.asarray(tensor, dtype=np.float64)))
        self.constraint = constraint

    @property
    def safe_value(self):
        return float(self.constraint(self._value))

    def __repr__(self):
        return f"ConstraintTrainVar({self._value!r}, constraint={self.constraint.__class__.__name__})"


class Module:
    def vars(self):
        out = {}
        for k, v in self.__dict__.items():
            if isinstance(v, TrainVar):
                out[k] = v
            elif isinstance(v, Module):
                out.update({f"{k}.{kk}": vv for kk, vv in v.vars().items()})
        return out
