"""Kernel factories with the signatures of the reference's experiments/nt_kernels.py:21-31, :83-103.

The reference returns a neural_tangents ``kernel_fn(x1, x2, get)`` closure; here the closure is a ``KernelFn``
object that carries the layer-stack description, so spax-level code can route whole objectives
(Gram + Cholesky + solve) through one fused CUDA call, while ``kernel_fn(x1, x2, get="nngp")`` itself still
works as a stand-alone Gram evaluation.  Conv / WideResNet kernels (nt_kernels.py:34-80) are out of scope.
"""
from __future__ import annotations

from . import device as _dev

__all__ = ["get_mlp_kernel", "get_dense_resnet_kernel", "KernelFn"]


class KernelFn:
    def __init__(self, spec: _dev.StackSpec, num_class, w_std, b_std, last_w_std):
        spec.ids()                                    # raises KeyError for an unsupported act (nt_kernels.py:18)
        self.spec = spec
        self.num_class = num_class
        self.w_std, self.b_std, self.last_w_std = float(w_std), float(b_std), float(last_w_std)

    def hp(self, device, eps=1e-6, alpha=2.0, beta=2.0):
        return _dev.make_hp(self.w_std, self.b_std, self.last_w_std, eps, alpha, beta, device=device)

    def hp_host(self, eps=1e-6, alpha=2.0, beta=2.0):
        return _dev._np_hp(self.w_std, self.b_std, self.last_w_std, eps, alpha, beta)

    def __call__(self, x1, x2=None, get="nngp"):
        if get not in ("nngp", ("nngp",)):
            raise NotImplementedError("only get='nngp' is on the accelerated path")
        x1 = _dev._f64(x1, "cuda" if not hasattr(x1, "device") else None)
        k = _dev.gram(x1, x2, spec=self.spec, hp=self.hp(x1.device))
        return k if get == "nngp" else (k,)


def get_mlp_kernel(num_hiddens, num_class=1, act="relu", w_std=1., b_std=0., last_w_std=1.):
    """(Dense(512, W_std, b_std), act) x num_hiddens, Dense(num_class, W_std=last_w_std)  - nt_kernels.py:21-31."""
    return KernelFn(_dev.StackSpec(num_hiddens, act, "mlp"), num_class, w_std, b_std, last_w_std)


def get_dense_resnet_kernel(num_hiddens, num_class=1, act="relu", w_std=1., b_std=0., last_w_std=1.):
    """Dense, num_hiddens x [z + Dense(act(z))], act, Dense(num_class, W_std=last_w_std) - nt_kernels.py:83-103."""
    return KernelFn(_dev.StackSpec(num_hiddens, act, "resnet"), num_class, w_std, b_std, last_w_std)
