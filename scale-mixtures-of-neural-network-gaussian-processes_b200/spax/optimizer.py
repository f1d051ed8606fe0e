"""Adam on the (host-resident, scalar) unconstrained variables of a spax Module - the role objax.optimizer.Adam plays
in the reference's train step (experiments/regression/train.py:61-67, :151).  The six trainable scalars live on the
host; the device work of a step is the single fused value+gradient call behind ``SPR.loss_and_grad``."""
import math

import numpy as np

__all__ = ["Adam"]


class Adam:
    def __init__(self, vc, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, skip_nan: bool = False):
        """vc: dict name -> TrainVar, as returned by ``Module.vars()`` (possibly filtered).
        skip_nan=True (a deviation from objax, opt-in): a step whose gradient holds a NaN (non-PD kernel matrix) is
        dropped instead of poisoning every variable."""
        self.vars = dict(vc)
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        self.skip_nan = skip_nan
        self.step = 0
        self.m = {k: np.zeros_like(v.value, dtype=np.float64) for k, v in self.vars.items()}
        self.v = {k: np.zeros_like(v.value, dtype=np.float64) for k, v in self.vars.items()}

    def __call__(self, lr: float, grads):
        """grads: dict name -> d loss / d (unconstrained value) (what ``SPR.loss_and_grad`` returns).  Variables
        without a gradient entry are left untouched.  The update is objax.optimizer.Adam's,
        ``p -= lr_t * m * rsqrt(v + eps)`` - eps sits INSIDE the square root, so a scalar whose gradient is ~1e-8
        (b_std at the reference defaults) moves by ~lr * 1e-4 per step instead of ~lr.  NaN gradients propagate
        into the variables like they do in the reference (the training loop stops on a NaN validation loss,
        regression/train.py:211) unless ``skip_nan`` was requested."""
        if self.skip_nan and any(math.isnan(float(g)) for g in grads.values()):
            return
        self.step += 1
        lr_t = lr * math.sqrt(1.0 - self.beta2 ** self.step) / (1.0 - self.beta1 ** self.step)
        for k, var in self.vars.items():
            if k not in grads:
                continue
            g = np.asarray(grads[k], dtype=np.float64)
            self.m[k] = self.beta1 * self.m[k] + (1.0 - self.beta1) * g
            self.v[k] = self.beta2 * self.v[k] + (1.0 - self.beta2) * g * g
            var.assign(var.value - lr_t * self.m[k] / np.sqrt(self.v[k] + self.eps))
