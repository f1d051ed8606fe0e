"""CPU oracle for the NNGP exact-GP hot path (NumPy / SciPy, FP64).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product path (the package next to
this directory) never does and fails loudly when its CUDA library is missing.

PARITY UNPINNED.  The reference (/root/reference) has no tests, golden vectors or fixtures for this path
and cannot be imported here (jax, neural_tangents and objax are absent from the image), so this file is a
*restatement* of the reference's algorithm, pinned instead against independent known answers
(``tests/golden/make_golden.py``: 50-digit mpmath evaluation, ``scipy.stats.multivariate_t``, analytic
identities, Monte-Carlo finite-width networks).

The arithmetic the reference delegates to third-party packages that are not vendored in its tree
(neural_tangents ~0.3.6-0.3.9 ``stax`` / ``predict``, jax ~0.2.2x ``lax.linalg`` / ``scipy.stats``; no lock
file exists, versions inferred from API usage - see SURVEY.md section 8c) is restated from their published
algorithms; every function cites the reference call site it follows.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla
from scipy.special import gammaln

__all__ = [
    "softplus", "softplus_inverse", "nngp_gram", "nngp_diag", "jitter", "multivariate_t_logpdf",
    "multivariate_normal_logpdf", "prior_logpdf", "spr_loss", "nt_predict", "student_t_logpdf",
    "normal_logpdf", "likelihood_logpdf", "spr_test_nll", "sample_f_iid_moments", "test_log_likelihood",
    "get_correct_count", "nngp_gram_dual", "spr_loss_grad", "find_grid_point",
]

ACTS = ("relu", "erf")
ARCHS = ("mlp", "resnet")


# ----------------------------------------------------------------------------------------------------------
# positive-constrained scalars: spax/bijectors.py:51-53 (Softplus.base / base_inv), spax/base.py:18-25
# ----------------------------------------------------------------------------------------------------------
def softplus(x):
    """``jax.nn.softplus`` = logaddexp(x, 0)  (spax/bijectors.py:52)."""
    return np.logaddexp(np.asarray(x, dtype=np.float64), 0.0)


def softplus_inverse(x):
    """spax/bijectors.py:53: where(x < 20, log(expm1(x)), x)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore"):
        return np.where(x < 20.0, np.log(np.expm1(x)), x)


# ----------------------------------------------------------------------------------------------------------
# NNGP recursion: experiments/nt_kernels.py:12-31 (get_mlp_kernel), :83-103 (get_dense_resnet_kernel)
# neural_tangents semantics restated: kernel_fn normalises X.X'^T by the feature count; Dense (NTK
# parameterisation) maps K -> W_std^2 K + b_std^2; Relu = ABRelu(0, 1); Erf closed form.
# ----------------------------------------------------------------------------------------------------------
def _act_diag(q, act):
    if act == "relu":
        return 0.5 * q                                              # (a^2+b^2)/2 * q with (a, b) = (0, 1)
    return (2.0 / math.pi) * np.arcsin(2.0 * q / (1.0 + 2.0 * q))   # stax.Erf, diagonal


def _act_offdiag(k, q1, q2, act):
    """One nonlinearity applied to the cross-covariance k [N,M] with marginal variances q1 [N], q2 [M]."""
    if act == "relu":
        prod = q1[:, None] * q2[None, :]
        s = np.sqrt(np.maximum(prod - k * k, 0.0))                  # NT _sqrt(., tol=0)
        theta = np.where((s == 0.0) & (k == 0.0), math.pi / 2, np.arctan2(s, k))  # NT _arctan2 fill pi/2
        return s / (2.0 * math.pi) + (0.5 - theta / (2.0 * math.pi)) * k
    prod = (1.0 + 2.0 * q1)[:, None] * (1.0 + 2.0 * q2)[None, :]
    return (2.0 / math.pi) * np.arcsin(2.0 * k / np.sqrt(prod))


def _recursion(k, q1, q2, *, num_hiddens, act, w_std, b_std, last_w_std, arch):
    """Shared layer stack.  ``k`` may be None (diagonal only).  Returns (k, q1, q2) after the last Dense."""
    if act not in ACTS:
        raise KeyError("Unsupported act '{}'".format(act))          # nt_kernels.py:18
    if arch not in ARCHS:
        raise ValueError(f"Unsupported network '{arch}'")           # regression/train.py:124
    w2, b2, v2 = w_std * w_std, b_std * b_std, last_w_std * last_w_std

    def dense(z):
        return None if z is None else w2 * z + b2

    if arch == "mlp":                                               # nt_kernels.py:25-28: (Dense, act) x L
        for _ in range(num_hiddens):
            k, q1, q2 = dense(k), dense(q1), dense(q2)
            k = None if k is None else _act_offdiag(k, q1, q2, act)
            q1, q2 = _act_diag(q1, act), _act_diag(q2, act)
    else:                                                           # nt_kernels.py:86-102
        k, q1, q2 = dense(k), dense(q1), dense(q2)                  # leading Dense(512)
        for _ in range(num_hiddens):                                # ResBlock: z + Dense(act(z))
            kb = None if k is None else dense(_act_offdiag(k, q1, q2, act))
            k = None if k is None else k + kb
            q1, q2 = q1 + dense(_act_diag(q1, act)), q2 + dense(_act_diag(q2, act))
        k = None if k is None else _act_offdiag(k, q1, q2, act)     # trailing act_class()
        q1, q2 = _act_diag(q1, act), _act_diag(q2, act)
    # final Dense(num_class, W_std=last_w_std), default b_std adds nothing (nt_kernels.py:29, :102)
    k = None if k is None else v2 * k
    return k, v2 * q1, v2 * q2


_CLIB = None


def _c_recursion():
    """Optional multi-threaded C restatement (oracle/nngp_recursion.c, built by oracle/Makefile)."""
    global _CLIB
    if _CLIB is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libnngp_recursion.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle`")
        lib = ctypes.CDLL(path)
        lib.nngp_recursion_inplace.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_double, ctypes.c_double, ctypes.c_double]
        _CLIB = lib
    return _CLIB


def nngp_gram(x1, x2=None, *, num_hiddens, act="relu", w_std=1.0, b_std=0.0, last_w_std=1.0, arch="mlp",
              row_block=4096, fast=False):
    """``kernel_fn(x1, x2, get="nngp")`` (spax/kernels.py:23-27) for the MLP / dense-resnet stacks.

    x1 [N,D], x2 [M,D] or None (=> x1).  Returns K [N,M] float64.  Evaluated in row blocks to bound memory.
    fast=True runs the layer recursion in the OpenMP C restatement (same arithmetic, all host cores) - used for
    the timed CPU baseline.
    """
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = x1 if x2 is None else np.ascontiguousarray(x2, dtype=np.float64)
    d = x1.shape[1]
    q1 = np.einsum("ij,ij->i", x1, x1) / d
    q2 = q1 if x2 is x1 else np.einsum("ij,ij->i", x2, x2) / d
    if fast:
        if act not in ACTS:
            raise KeyError("Unsupported act '{}'".format(act))
        if arch not in ARCHS:
            raise ValueError(f"Unsupported network '{arch}'")
        # X1.X2^T / D in row blocks written straight into the output: one big `x1 @ x2.T` segfaults inside NumPy /
        # OpenBLAS once the result has more than 2^31 elements (N >= 46 341), and the blocks avoid a second N x M array
        k = np.empty((x1.shape[0], x2.shape[0]), dtype=np.float64)
        x2t = x2.T
        for r0 in range(0, x1.shape[0], row_block):
            r1 = min(r0 + row_block, x1.shape[0])
            np.matmul(x1[r0:r1], x2t, out=k[r0:r1])
        k /= d
        rc = _c_recursion().nngp_recursion_inplace(k.ctypes.data, k.shape[0], k.shape[1], q1.ctypes.data,
                                                   q2.ctypes.data, int(num_hiddens), ACTS.index(act),
                                                   ARCHS.index(arch), float(w_std), float(b_std), float(last_w_std))
        if rc != 0:
            raise MemoryError("nngp_recursion_inplace")
        return k
    out = np.empty((x1.shape[0], x2.shape[0]), dtype=np.float64)
    kw = dict(num_hiddens=num_hiddens, act=act, w_std=w_std, b_std=b_std, last_w_std=last_w_std, arch=arch)
    for r0 in range(0, x1.shape[0], row_block):
        r1 = min(r0 + row_block, x1.shape[0])
        k0 = (x1[r0:r1] @ x2.T) / d
        out[r0:r1], _, _ = _recursion(k0, q1[r0:r1], q2, **kw)
    return out


def nngp_diag(x, *, num_hiddens, act="relu", w_std=1.0, b_std=0.0, last_w_std=1.0, arch="mlp"):
    """Marginal variances q_i of the same stack (what NT carries as cov1 / cov2)."""
    x = np.asarray(x, dtype=np.float64)
    q = np.einsum("ij,ij->i", x, x) / x.shape[1]
    _, q, _ = _recursion(None, q, q, num_hiddens=num_hiddens, act=act, w_std=w_std, b_std=b_std,
                         last_w_std=last_w_std, arch=arch)
    return q


def jitter(num, eps=1e-6):
    """spax/utils.py:26-27."""
    return eps * np.eye(num)


# ----------------------------------------------------------------------------------------------------------
# marginal likelihoods: spax/utils.py:160-183, spax/likelihoods.py:25-28, :45-50, spax/models.py:93-98
# ----------------------------------------------------------------------------------------------------------
# SciPy's LAPACK interface uses 32-bit integers: one potrf / trtrs call on a matrix with more than 2^31 elements
# (N >= 46 341) segfaults.  Above this order the factorisation and the forward substitution run block by block
# (right-looking, LAPACK potrf on the diagonal block, trsm for the panel, gemm for the trailing blocks): the same LAPACK /
# BLAS family, the same arithmetic up to the usual blocked-algorithm rounding (checked in tests/test_oracle_cpu.py).
BLOCKED_ABOVE = 40000


def cholesky_blocked_inplace(a, nb=8192):
    """Lower Cholesky factor of the symmetric a (C-order, lower part read) IN PLACE, nb-wide panels."""
    n = a.shape[0]
    for k0 in range(0, n, nb):
        k1 = min(k0 + nb, n)
        lkk = sla.cholesky(a[k0:k1, k0:k1], lower=True, check_finite=False)
        a[k0:k1, k0:k1] = lkk
        if k1 == n:
            break
        panel = a[k1:, k0:k1]
        panel[...] = sla.solve_triangular(lkk, panel.T, lower=True, check_finite=False).T      # panel L_kk^-T
        for i0 in range(k1, n, nb):
            i1 = min(i0 + nb, n)
            a[i0:i1, k1:i1] -= panel[i0 - k1:i1 - k1] @ panel[:i1 - k1].T
    return a


def forward_substitution_blocked(L, y, nb=8192):
    """z = L^-1 y with the lower factor stored in the lower part of L (blocks of nb rows)."""
    n = L.shape[0]
    z = np.array(y, dtype=np.float64)
    for k0 in range(0, n, nb):
        k1 = min(k0 + nb, n)
        if k0 > 0:
            z[k0:k1] -= L[k0:k1, :k0] @ z[:k0]
        z[k0:k1] = sla.solve_triangular(L[k0:k1, k0:k1], z[k0:k1], lower=True, check_finite=False)
    return z


def _chol_and_solve(shape, rhs):
    """(L, L^-1 rhs) for the SPD matrix `shape`; raises scipy.linalg.LinAlgError when it is not positive definite.
    Large matrices are factored block-wise IN PLACE (`shape` is overwritten: callers pass a temporary)."""
    if shape.shape[0] > BLOCKED_ABOVE:
        L = cholesky_blocked_inplace(shape)
        return L, forward_substitution_blocked(L, rhs)
    L = sla.cholesky(shape, lower=True, check_finite=False)
    return L, sla.solve_triangular(L, rhs, lower=True, check_finite=False)


def multivariate_t_logpdf(x, loc, shape, df):
    """spax/utils.py:178-183: Cholesky, L^-1 (x - loc), closed form."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    t = 0.5 * (df + n)
    try:
        L, y = _chol_and_solve(shape, x - loc)
    except sla.LinAlgError:
        return float("nan")                                         # jax cholesky: NaN on non-PD, no raise
    return float(-t * np.log(1.0 + (1.0 / df) * (y @ y)) - n / 2 * np.log(df * np.pi) + gammaln(t)
                 - gammaln(0.5 * df) - np.log(np.diag(L)).sum())


def multivariate_normal_logpdf(x, mean, cov):
    """``jax.scipy.stats.multivariate_normal.logpdf`` as called at spax/likelihoods.py:27."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    try:
        L, y = _chol_and_solve(cov, x - mean)
    except sla.LinAlgError:
        return float("nan")
    return float(-0.5 * (y @ y) - n / 2 * np.log(2 * np.pi) - np.log(np.diag(L)).sum())


def prior_logpdf(y, cov, *, kind, a=None, b=None):
    """Likelihood.prior_logpdf: kind 'student_t' (likelihoods.py:45-50) or 'gauss' (:25-28)."""
    zero = np.zeros_like(y)
    if kind == "student_t":
        if cov.shape[0] > BLOCKED_ABOVE:      # scale in place: a second N x N array would not fit next to the first
            cov *= b / a
            return multivariate_t_logpdf(y, zero, cov, 2 * a)
        return multivariate_t_logpdf(y, zero, (b / a) * cov, 2 * a)
    if kind == "gauss":
        return multivariate_normal_logpdf(y, zero, cov)
    raise KeyError(kind)


def spr_loss(x, y, *, num_hiddens, act="relu", w_std=1.0, b_std=0.0, last_w_std=1.0, arch="mlp", eps=1e-6,
             kind="student_t", a=2.0, b=2.0, fast=False):
    """SPR.loss (spax/models.py:93-98): -prior_logpdf(y, K + eps I) / N."""
    n = x.shape[0]
    cov = nngp_gram(x, num_hiddens=num_hiddens, act=act, w_std=w_std, b_std=b_std, last_w_std=last_w_std,
                    arch=arch, fast=fast)
    cov[np.diag_indices(n)] += eps                                  # == + jitter(n, eps) without the N x N eye
    return -prior_logpdf(y, cov, kind=kind, a=a, b=b) / n


# ----------------------------------------------------------------------------------------------------------
# gradient of SPR.loss w.r.t. the six positive scalars - what objax.GradValues(model.loss, vars) computes in the
# training loop (experiments/regression/train.py:62-66, :178-179) by reverse-mode AD through the same path.
# Restated analytically: forward-mode duals of the layer recursion + d log p / d Sigma in closed form.
# ----------------------------------------------------------------------------------------------------------
def _act_diag_dual(q, act):
    """(phi(q), dphi/dq) on the diagonal."""
    if act == "relu":
        return 0.5 * q, np.full_like(q, 0.5)
    return ((2.0 / math.pi) * np.arcsin(2.0 * q / (1.0 + 2.0 * q)),
            (4.0 / math.pi) / ((1.0 + 2.0 * q) * np.sqrt(1.0 + 4.0 * q)))


def _act_offdiag_dual(k, q1, q2, act):
    """(phi, dphi/dk, dphi/dq1, dphi/dq2), all [N,M], for one nonlinearity."""
    if act == "relu":
        prod = q1[:, None] * q2[None, :]
        s = np.sqrt(np.maximum(prod - k * k, 0.0))
        theta = np.where((s == 0.0) & (k == 0.0), math.pi / 2, np.arctan2(s, k))
        phi = s / (2.0 * math.pi) + (0.5 - theta / (2.0 * math.pi)) * k
        with np.errstate(divide="ignore", invalid="ignore"):
            i1 = np.where(q1 > 0.0, 1.0 / (4.0 * math.pi * q1), 0.0)
            i2 = np.where(q2 > 0.0, 1.0 / (4.0 * math.pi * q2), 0.0)
        return phi, 0.5 - theta / (2.0 * math.pi), s * i1[:, None], s * i2[None, :]
    t1, t2 = 1.0 / np.sqrt(1.0 + 2.0 * q1), 1.0 / np.sqrt(1.0 + 2.0 * q2)
    x = 2.0 * k * t1[:, None] * t2[None, :]
    g = (2.0 / math.pi) / np.sqrt(np.maximum(1.0 - x * x, 1e-300))
    phi = (2.0 / math.pi) * np.arcsin(np.clip(x, -1.0, 1.0))
    # d t / d q = -t^3
    return (phi, g * 2.0 * t1[:, None] * t2[None, :], g * (-x) * (t1 * t1)[:, None], g * (-x) * (t2 * t2)[None, :])


def nngp_gram_dual(x1, x2=None, *, num_hiddens, act="relu", w_std=1.0, b_std=0.0, last_w_std=1.0, arch="mlp"):
    """K and its partial derivatives w.r.t. (w_std, b_std, last_w_std): returns (K, dK_w, dK_b, dK_v)."""
    if act not in ACTS:
        raise KeyError("Unsupported act '{}'".format(act))
    if arch not in ARCHS:
        raise ValueError(f"Unsupported network '{arch}'")
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = x1 if x2 is None else np.ascontiguousarray(x2, dtype=np.float64)
    d = x1.shape[1]
    w2, b2, v2 = w_std * w_std, b_std * b_std, last_w_std * last_w_std

    # state: value + (d/dw, d/db) for k [N,M], q1 [N], q2 [M]
    k = (x1 @ x2.T) / d
    q1 = np.einsum("ij,ij->i", x1, x1) / d
    q2 = np.einsum("ij,ij->i", x2, x2) / d
    zk, z1, z2 = np.zeros_like(k), np.zeros_like(q1), np.zeros_like(q2)
    K, Q1, Q2 = (k, zk, zk), (q1, z1, z1), (q2, z2, z2)

    def dense(t):
        v, dw, db = t
        return (w2 * v + b2, 2.0 * w_std * v + w2 * dw, w2 * db + 2.0 * b_std)

    def act_k(K, Q1, Q2):
        phi, pk, p1, p2 = _act_offdiag_dual(K[0], Q1[0], Q2[0], act)
        return (phi,) + tuple(pk * K[i] + p1 * Q1[i][:, None] + p2 * Q2[i][None, :] for i in (1, 2))

    def act_q(Q):
        phi, pq = _act_diag_dual(Q[0], act)
        return (phi, pq * Q[1], pq * Q[2])

    def add(a, b):
        return tuple(u + v for u, v in zip(a, b))

    if arch == "mlp":
        for _ in range(num_hiddens):
            K, Q1, Q2 = dense(K), dense(Q1), dense(Q2)
            K, Q1, Q2 = act_k(K, Q1, Q2), act_q(Q1), act_q(Q2)
    else:
        K, Q1, Q2 = dense(K), dense(Q1), dense(Q2)
        for _ in range(num_hiddens):
            K, Q1, Q2 = add(K, dense(act_k(K, Q1, Q2))), add(Q1, dense(act_q(Q1))), add(Q2, dense(act_q(Q2)))
        K = act_k(K, Q1, Q2)
    return v2 * K[0], v2 * K[1], v2 * K[2], 2.0 * last_w_std * K[0]


def spr_loss_grad(x, y, *, num_hiddens, act="relu", w_std=1.0, b_std=0.0, last_w_std=1.0, arch="mlp", eps=1e-6,
                  kind="student_t", a=2.0, b=2.0):
    """(loss, grad) with grad = d loss / d (w_std, b_std, last_w_std, eps, alpha, beta), loss = SPR.loss
    (spax/models.py:93-98).  With A = K + eps I, alpha_v = A^-1 y, quad = y^T A^-1 y:
        d log p / d A = 1/2 (gamma alpha_v alpha_v^T - A^-1),  gamma = (2a + N) / ((2a + a quad / b) b / a)  [t],  1 [gauss]
    and the (a, b) dependence of the Student-t density in closed form (quad / nu = quad / (2b) does not depend on a)."""
    from scipy.special import digamma
    n = x.shape[0]
    K, dKw, dKb, dKv = nngp_gram_dual(x, num_hiddens=num_hiddens, act=act, w_std=w_std, b_std=b_std,
                                      last_w_std=last_w_std, arch=arch)
    A = K + eps * np.eye(n)
    try:
        c = sla.cho_factor(A, lower=True, check_finite=False)
    except sla.LinAlgError:
        return float("nan"), np.full(6, np.nan)
    Ainv = sla.cho_solve(c, np.eye(n), check_finite=False)
    al = Ainv @ y
    quad = float(y @ al)
    logdet_half = float(np.log(np.diag(c[0])).sum())
    if kind == "student_t":
        nu, t = 2.0 * a, a + 0.5 * n
        logp = (-t * np.log1p(quad / (2.0 * b)) - 0.5 * n * np.log(nu * np.pi) + gammaln(t) - gammaln(a)
                - 0.5 * n * np.log(b / a) - logdet_half)
        gamma = (nu + n) / ((nu + quad * a / b) * (b / a))
        dlp_da = -np.log1p(quad / (2.0 * b)) + digamma(t) - digamma(a)
        dlp_db = t * quad / (b * (2.0 * b + quad)) - 0.5 * n / b
    elif kind == "gauss":
        logp = -0.5 * quad - 0.5 * n * np.log(2 * np.pi) - logdet_half
        gamma, dlp_da, dlp_db = 1.0, 0.0, 0.0
    else:
        raise KeyError(kind)
    G = 0.5 * (gamma * np.outer(al, al) - Ainv)
    dlp = np.array([np.sum(G * dKw), np.sum(G * dKb), np.sum(G * dKv), np.trace(G), dlp_da, dlp_db])
    return float(-logp / n), -dlp / n


# ----------------------------------------------------------------------------------------------------------
# predictive: spax/kernels.py:29-32 -> neural_tangents.predict.gradient_descent_mse_ensemble (t=None,
# get="nngp", compute_cov=True, diag_reg_absolute_scale=False)
# ----------------------------------------------------------------------------------------------------------
def nt_predict(x, y, x_test, eps=1e-6, *, kernel_kwargs, k_dd=None):
    """Returns (mean [T,C], cov [T,T]).  y is [N,C].  The regulariser is RELATIVE: eps * tr(K_dd) / N."""
    y = np.asarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None]
    n = x.shape[0]
    if k_dd is None:
        k_dd = nngp_gram(x, **kernel_kwargs)
    k_td = nngp_gram(x_test, x, **kernel_kwargs)
    k_tt = nngp_gram(x_test, **kernel_kwargs)
    reg = eps * np.trace(k_dd) / n
    a1 = k_dd + reg * np.eye(n)
    try:
        c = sla.cho_factor(a1, lower=True, check_finite=False)
    except sla.LinAlgError:
        t = x_test.shape[0]
        return np.full((t, y.shape[1]), np.nan), np.full((t, t), np.nan)
    mean = k_td @ sla.cho_solve(c, y, check_finite=False)
    cov = k_tt - k_td @ sla.cho_solve(c, k_td.T, check_finite=False)
    return mean, cov


def find_grid_point(x, y, x_test, eps, *, kernel_kwargs):
    """One (w_std, b_std, eps) point of experiments/regression/find.py:134-160: predict(eps) (relative regulariser,
    :75-77) -> (mean [T], diag cov [T]); log det(K + eps I) and y^T (K + eps I)^-1 y with the ABSOLUTE jitter
    (:149-156; the reference forms the explicit inverse, the value is the same)."""
    n = x.shape[0]
    k_dd = nngp_gram(x, **kernel_kwargs)
    mean, cov = nt_predict(x, y, x_test, eps, kernel_kwargs=kernel_kwargs, k_dd=k_dd)
    a = k_dd + eps * np.eye(n)
    sign, logdet = np.linalg.slogdet(a)
    quad = float(y @ np.linalg.solve(a, y))
    return mean.ravel(), np.diag(cov).copy(), float(logdet), quad


def _t_logpdf(x, df, loc, scale):
    """``jax.scipy.stats.t.logpdf`` (likelihoods.py:64)."""
    z = (x - loc) / scale
    norm = gammaln(df / 2) + 0.5 * np.log(df) + 0.5 * np.log(scale * scale * np.pi) - gammaln((df + 1) / 2)
    return -(norm + (df + 1) / 2 * np.log1p(z * z / df))


def student_t_logpdf(x, mean, cov, cov_data, y_data, *, a, b):
    """StudentTLikelihood.logpdf (spax/likelihoods.py:52-65); uses the explicit inverse like the reference."""
    n = cov_data.shape[-1]
    df = 2 * a
    cond_df = df + n
    inv_cov_data = np.linalg.inv(b / a * cov_data + jitter(n))      # default jitter 1e-6 (likelihoods.py:60)
    d = df + y_data @ inv_cov_data @ y_data
    sigma = np.sqrt(np.diag(d / cond_df * b / a * cov))
    return _t_logpdf(x, cond_df, mean, sigma)


def normal_logpdf(x, mean, cov):
    """GaussianLikelihood.logpdf (spax/likelihoods.py:30-33)."""
    sigma = np.sqrt(np.diag(cov))
    return -0.5 * np.log(2 * np.pi) - np.log(sigma) - 0.5 * ((x - mean) / sigma) ** 2


def likelihood_logpdf(x, mean, cov, aux, *, kind, a=None, b=None):
    if kind == "student_t":
        return student_t_logpdf(x, mean, cov, aux[0], aux[1], a=a, b=b)
    return normal_logpdf(x, mean, cov)


def spr_test_nll(x_data, y_data, x_test, y_test, y_mean, y_std, *, num_hiddens, act="relu", w_std=1.0,
                 b_std=0.0, last_w_std=1.0, arch="mlp", eps=1e-6, kind="student_t", a=2.0, b=2.0,
                 return_parts=False):
    """SPR.test_nll (spax/models.py:100-120), including its three different regularisers."""
    kw = dict(num_hiddens=num_hiddens, act=act, w_std=w_std, b_std=b_std, last_w_std=last_w_std, arch=arch)
    cov_data = nngp_gram(x_data, **kw)                              # models.py:107 (no jitter: "TODO: check")
    mean, cov = nt_predict(x_data, y_data[:, None], x_test, eps, kernel_kwargs=kw, k_dd=cov_data)
    aux = (cov_data, y_data) if kind == "student_t" else None
    log_prob = likelihood_logpdf(y_test * y_std + y_mean, mean.ravel() * y_std + y_mean, cov * y_std ** 2,
                                 aux, kind=kind, a=a, b=b)
    nll = -float(np.mean(log_prob))
    if return_parts:
        return nll, mean.ravel(), np.diag(cov).copy()
    return nll


def sample_f_iid_moments(mean, cov, *, a, b):
    """InverseGammaPrior.sample_f_iid (spax/priors.py:60-68): f = mean + sigma * t_{2a}, sigma from the
    diagonal of (b/a) cov.  Returns (loc, sigma, df) - the distribution, not draws."""
    sigma = np.sqrt(np.diagonal(b / a * cov, axis1=-2, axis2=-1))
    return mean, sigma, 2 * a


# ----------------------------------------------------------------------------------------------------------
# Monte-Carlo classification metrics on draws sampled_f [C, B, S]: spax/utils.py:47-74
# ----------------------------------------------------------------------------------------------------------
def _log_softmax0(f):
    m = f.max(axis=0, keepdims=True)
    return f - (m + np.log(np.exp(f - m).sum(axis=0, keepdims=True)))


def _logsumexp(x, axis):
    m = x.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(x - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def test_log_likelihood(sampled_f, label):
    """spax/utils.py:61-66."""
    num_samples = sampled_f.shape[2]
    lsm = _log_softmax0(sampled_f)                                       # [C, B, S]
    true = np.take_along_axis(lsm, np.asarray(label)[None, :, None].repeat(num_samples, axis=2), axis=0)[0]
    return float(np.mean(_logsumexp(true, 1) - np.log(num_samples)))


def get_correct_count(sampled_f, label):
    """spax/utils.py:69-74."""
    lsm = _log_softmax0(sampled_f)
    y_pred = np.argmax(_logsumexp(lsm, 2), axis=0)
    return int(np.sum(y_pred == np.asarray(label)))
