"""NNGPKernel with the interface of spax/kernels.py:9-41."""
from .base import Module, ConstraintTrainVar
from .bijectors import positive
from .. import device as _dev
from ..nt_kernels import KernelFn

__all__ = ["NNGPKernel"]


class NNGPKernel(Module):
    def __init__(self, get_kernel_fn, w_std: float = 1.0, b_std: float = 1.0, last_w_std: float = 1.0):
        self._get_kernel_fn = get_kernel_fn
        self.w_std = ConstraintTrainVar(w_std, constraint=positive())
        self.b_std = ConstraintTrainVar(b_std, constraint=positive())
        self.last_w_std = ConstraintTrainVar(last_w_std, constraint=positive())

    def K(self, kernel_fn, x, x2=None):
        """spax/kernels.py:23-27.  x2 None / same object -> symmetric path (lower tiles + mirror)."""
        return kernel_fn(x, None if (x2 is None or x2 is x) else x2, get="nngp")

    def predict(self, kernel_fn, x, y, x_test, eps=1e-6, full_cov=True):
        """spax/kernels.py:29-32 (neural_tangents gradient_descent_mse_ensemble(..., diag_reg=eps)(x_test, "nngp",
        compute_cov=True), relative regulariser).  Returns ``(mean [T, C], cov [T, T])`` exactly like the reference
        (its callers do ``cov * y_std ** 2`` and take the diagonal: spax/models.py:116, regression/find.py:146).
        ``full_cov=False`` is the cheap variant for callers that only need diag(cov): ``(mean, var [T])``.
        NumPy inputs go through the device path too and come back as NumPy arrays."""
        if not isinstance(kernel_fn, KernelFn):
            raise TypeError("predict needs a kernel_fn built by smnngp nt_kernels")
        import numpy as np
        host = isinstance(x, np.ndarray)
        if host and not full_cov:
            mean, var, _ = _dev.predict(x, y, x_test, spec=kernel_fn.spec, hp=kernel_fn.hp_host(eps=eps))
            return mean, var
        if host:
            import torch
            dev = torch.device("cuda", torch.cuda.current_device())
            x, y, x_test = (torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64), device=dev) for v in (x, y, x_test))
        mean, second, _ = _dev.predict(x, y, x_test, spec=kernel_fn.spec, hp=kernel_fn.hp(x.device, eps=eps),
                                       full_cov=full_cov)
        if host:
            return mean.cpu().numpy(), second.cpu().numpy()
        return mean, second

    def get_params(self):
        return (self.w_std.safe_value, self.b_std.safe_value, self.last_w_std.safe_value)

    def get_kernel_fn(self):
        return self._get_kernel_fn(self.w_std.safe_value, self.b_std.safe_value, self.last_w_std.safe_value)
