"""Hot subset of spax/utils.py: jitter (:26-27) and multivariate_t_logpdf (:160-183) on the CUDA path."""
import math

import torch

from .. import device as _dev

__all__ = ["jitter", "multivariate_t_logpdf", "multivariate_normal_logpdf"]


def jitter(num, eps=1e-6, device="cuda"):
    """eps * I.  The fused entry points fold this into the Gram epilogue instead of materialising it."""
    return eps * torch.eye(num, dtype=torch.float64, device=device)


def multivariate_t_logpdf(x, loc, shape, df):
    """spax/utils.py:178-183 with chol / triangular solve / reductions on the GPU."""
    x = _dev._f64(x, shape.device) - loc
    n = x.shape[-1]
    logdet, quad, info = _dev.cov_solve(shape, x)
    t = 0.5 * (df + n)
    return (-t * torch.log(1 + quad / df) - n / 2 * math.log(df * math.pi) + math.lgamma(t)
            - math.lgamma(0.5 * df) - logdet)


def multivariate_normal_logpdf(x, mean, cov):
    """jax.scipy.stats.multivariate_normal.logpdf as used at spax/likelihoods.py:27."""
    x = _dev._f64(x, cov.device) - mean
    n = x.shape[-1]
    logdet, quad, info = _dev.cov_solve(cov, x)
    return -0.5 * quad - n / 2 * math.log(2 * math.pi) - logdet
