"""Prior draw stage with the interface of spax/priors.py (sample_f_iid only - the part of the priors that the
exact-GP hot path touches; sample_f / kl_divergence belong to the sparse variational path, out of scope)."""
from .base import Module, ConstraintTrainVar
from .bijectors import positive
from .. import device as _dev

__all__ = ["Prior", "GaussianPrior", "InverseGammaPrior"]


class Prior(Module):
    kind = None

    def _hp(self, device):
        a = self.a.safe_value if hasattr(self, "a") else 2.0
        b = self.b.safe_value if hasattr(self, "b") else 2.0
        return _dev.make_hp(1.0, 0.0, 1.0, 1e-6, a, b, device=device)

    def sample_f_iid(self, key, mean, cov, num_samples):
        """mean [C, B] (the reference passes mean.T), cov [C, B, B] / [B, B] or its diagonal [C, B] / [B];
        key = integer seed.  Returns [C, B, S]."""
        var = cov
        if cov.ndim == 3 or (cov.ndim == 2 and cov.shape[-1] == cov.shape[-2] and cov.shape[0] != mean.shape[0]):
            var = cov.diagonal(dim1=-2, dim2=-1)
        elif cov.ndim == 2 and cov.shape == (mean.shape[1], mean.shape[1]):
            var = cov.diagonal()
        return _dev.sample_f_iid(mean.T.contiguous(), var.contiguous(), hp=self._hp(mean.device), kind=self.kind,
                                 num_samples=num_samples, seed=int(key))


class GaussianPrior(Prior):
    kind = "gauss"


class InverseGammaPrior(Prior):
    kind = "student_t"

    def __init__(self, alpha, beta):
        self.alpha, self.beta = alpha, beta
        self.a = ConstraintTrainVar(alpha, constraint=positive())
        self.b = ConstraintTrainVar(beta, constraint=positive())
