"""SPR with the interface of spax/models.py:81-120.  When the kernel closure is an smnngp ``KernelFn`` and the
likelihood is one of ours, ``loss`` and ``test_nll`` are single fused C-ABI calls (Gram never leaves the
device, solves fused into the factorisation); otherwise they compose K / prior_logpdf / predict / logpdf
exactly like the reference does - every piece still on the CUDA path."""
import numpy as np
import torch

from .base import Module, ConstraintTrainVar
from .bijectors import positive
from .utils import jitter
from .. import device as _dev
from ..nt_kernels import KernelFn

__all__ = ["SPR", "DistributedSPR"]


class SPR(Module):
    def __init__(self, kernel, likelihood, x_data, y_data, y_mean, y_std, *, eps: float = 1e-6):
        self.kernel = kernel
        self.likelihood = likelihood
        self.x_data = x_data
        self.y_data = y_data
        self.y_mean = y_mean
        self.y_std = y_std
        self.num_data = x_data.shape[0]
        self.eps = ConstraintTrainVar(eps, constraint=positive())

    def _hp_args(self):
        lik = self.likelihood
        a = lik.a.safe_value if hasattr(lik, "a") else 2.0
        b = lik.b.safe_value if hasattr(lik, "b") else 2.0
        return dict(eps=self.eps.safe_value, alpha=a, beta=b)

    def _fused(self, kernel_fn):
        return isinstance(kernel_fn, KernelFn) and getattr(self.likelihood, "kind", None) in _dev.KIND

    def loss(self):
        kernel_fn = self.kernel.get_kernel_fn()
        if self._fused(kernel_fn):
            host = isinstance(self.x_data, np.ndarray)
            hp = kernel_fn.hp_host(**self._hp_args()) if host else kernel_fn.hp(self.x_data.device, **self._hp_args())
            out, _ = _dev.lml(self.x_data, self.y_data, spec=kernel_fn.spec, hp=hp, kind=self.likelihood.kind)
            return float(out[1]) if host else out[1]
        eps = self.eps.safe_value
        cov = self.kernel.K(kernel_fn, self.x_data) + jitter(self.num_data, eps=eps, device=self.x_data.device)
        log_prob = self.likelihood.prior_logpdf(self.y_data, cov)
        return -log_prob / self.num_data

    def loss_and_grad(self):
        """(loss, grads): the pair ``objax.GradValues(model.loss, model.vars())`` returns in the reference's train
        step (experiments/regression/train.py:62-66), as one fused call.  ``grads`` maps the names of
        ``self.vars()`` to d loss / d (unconstrained variable value) - the softplus chain rule
        (spax/base.py:23-25, spax/bijectors.py:51-53) is applied here.  Scalars the likelihood does not own
        (Gaussian: alpha, beta) are absent."""
        kernel_fn = self.kernel.get_kernel_fn()
        if not self._fused(kernel_fn):
            raise TypeError("loss_and_grad needs a kernel_fn built by smnngp nt_kernels and a spax likelihood")
        host = isinstance(self.x_data, np.ndarray)
        hp = kernel_fn.hp_host(**self._hp_args()) if host else kernel_fn.hp(self.x_data.device, **self._hp_args())
        out, grad, _ = _dev.lml_grad(self.x_data, self.y_data, spec=kernel_fn.spec, hp=hp, kind=self.likelihood.kind)
        g = grad if host else grad.cpu().numpy()
        loss = float(out[1])
        slots = {"kernel.w_std": (self.kernel.w_std, 0), "kernel.b_std": (self.kernel.b_std, 1),
                 "kernel.last_w_std": (self.kernel.last_w_std, 2), "eps": (self.eps, 3)}
        if hasattr(self.likelihood, "a"):
            slots["likelihood.a"] = (self.likelihood.a, 4)
            slots["likelihood.b"] = (self.likelihood.b, 5)
        grads = {name: float(g[i]) * float(var.constraint.grad(var.value)) for name, (var, i) in slots.items()}
        return loss, grads

    def test_nll(self, x, y):
        kernel_fn = self.kernel.get_kernel_fn()
        if self._fused(kernel_fn):
            host = isinstance(self.x_data, np.ndarray)
            hp = kernel_fn.hp_host(**self._hp_args()) if host else kernel_fn.hp(self.x_data.device, **self._hp_args())
            nll, _, _, _ = _dev.test_nll(self.x_data, self.y_data, x, y, float(self.y_mean), float(self.y_std),
                                         spec=kernel_fn.spec, hp=hp, kind=self.likelihood.kind)
            return nll
        eps = self.eps.safe_value
        mean, cov = self.kernel.predict(kernel_fn, self.x_data, self.y_data[:, None], x, eps=eps)
        require = self.likelihood.require
        if require:
            aux_dict = dict(y_data=self.y_data)
            if "cov_data" in require:
                aux_dict["cov_data"] = self.kernel.K(kernel_fn, self.x_data)      # models.py:107, no jitter
            aux = tuple(aux_dict[k] for k in require)
        else:
            aux = None
        log_prob = self.likelihood.logpdf((y * self.y_std) + self.y_mean, (mean.flatten() * self.y_std) + self.y_mean,
                                          cov * self.y_std ** 2, aux)
        return -torch.mean(log_prob)



class DistributedSPR(SPR):
    """SPR whose device work is sharded over the ranks of a torch.distributed process group (one process per GPU,
    launched with torchrun): the same methods, every rank passes the same data and obtains the same numbers.

        loss / loss_and_grad -> smnngp_lml_grad_mg_f64  (distributed.DistributedGrad)
        test_nll             -> smnngp_test_nll_mg_f64  (distributed.DistributedPredict)

    x_data / y_data must be torch CUDA tensors on this rank's device.  The solvers are created on first use (they own
    the rank's row shard and the NVLink-visible buffers) and re-used by every later step - the reference evaluates the
    same shapes for 30 000 steps (experiments/regression/train.py:178).  ``solvers=`` injects ready-made ones (tests)."""

    def __init__(self, kernel, likelihood, x_data, y_data, y_mean, y_std, *, eps: float = 1e-6, group=None, block=None,
                 solvers=None):
        super().__init__(kernel, likelihood, x_data, y_data, y_mean, y_std, eps=eps)
        self.group, self.block = group, block
        self._grad_solver = (solvers or {}).get("grad")
        self._predict_solvers = dict((solvers or {}).get("predict", {}))

    def _spec_hp(self):
        kernel_fn = self.kernel.get_kernel_fn()
        if not self._fused(kernel_fn):
            raise TypeError("DistributedSPR needs a kernel_fn built by smnngp nt_kernels and a spax likelihood")
        return kernel_fn.spec, kernel_fn.hp(self.x_data.device, **self._hp_args())

    def loss_and_grad(self):
        from ..distributed import DistributedGrad
        spec, hp = self._spec_hp()
        if self._grad_solver is None:
            self._grad_solver = DistributedGrad(self.num_data, self.x_data.shape[1], spec, self.x_data.device,
                                                group=self.group, block=self.block)
        out, grad, _ = self._grad_solver.lml_grad(self.x_data, self.y_data, hp, kind=self.likelihood.kind)
        g = grad.cpu().numpy()
        slots = {"kernel.w_std": (self.kernel.w_std, 0), "kernel.b_std": (self.kernel.b_std, 1),
                 "kernel.last_w_std": (self.kernel.last_w_std, 2), "eps": (self.eps, 3)}
        if hasattr(self.likelihood, "a"):
            slots["likelihood.a"] = (self.likelihood.a, 4)
            slots["likelihood.b"] = (self.likelihood.b, 5)
        grads = {name: float(g[i]) * float(var.constraint.grad(var.value)) for name, (var, i) in slots.items()}
        return float(out[1]), grads

    def loss(self):
        return self.loss_and_grad()[0]

    def test_nll(self, x, y):
        from ..distributed import DistributedPredict
        spec, hp = self._spec_hp()
        t = int(x.shape[0])
        solver = self._predict_solvers.get(t)
        if solver is None:
            solver = DistributedPredict(self.num_data, self.x_data.shape[1], t, 1, spec, self.x_data.device,
                                        group=self.group, block=self.block)
            self._predict_solvers[t] = solver
        nll, _, _, _ = solver.test_nll(self.x_data, self.y_data, x, y, float(self.y_mean), float(self.y_std), hp,
                                       kind=self.likelihood.kind)
        return nll[0]

    def close(self):
        for s in [self._grad_solver, *self._predict_solvers.values()]:
            if s is not None and hasattr(s, "close"):
                s.close()
        self._grad_solver, self._predict_solvers = None, {}
