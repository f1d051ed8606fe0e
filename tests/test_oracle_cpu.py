"""CPU suite (-m "not gpu"): pins the oracle against independent known answers and checks the host-side pieces
that need no GPU (C-ABI exports, mirrored spax scalars)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import scipy.stats

from oracle import nngp_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "nngp_golden.npz"))


def _variants():
    for vi, s in enumerate(GOLD["variants"]):
        act, arch, L, w, b, v = str(s).split(",")
        yield vi, dict(num_hiddens=int(L), act=act, arch=arch, w_std=float(w), b_std=float(b), last_w_std=float(v))


@pytest.mark.parametrize("vi,kw", list(_variants()))
def test_oracle_matches_mpmath_golden(vi, kw):
    x, xt, y, yt = GOLD["x"], GOLD["xt"], GOLD["y"], GOLD["yt"]
    eps, a, b = float(GOLD["eps"]), float(GOLD["a"]), float(GOLD["b"])
    ym, ys = float(GOLD["y_mean"]), float(GOLD["y_std"])
    K = orc.nngp_gram(x, **kw)
    assert np.abs(K - GOLD[f"K{vi}"]).max() <= 1e-13 * np.abs(K).max()
    Ktd = orc.nngp_gram(xt, x, **kw)
    assert np.abs(Ktd - GOLD[f"Ktd{vi}"]).max() <= 1e-13 * np.abs(K).max()
    for kind, key in (("student_t", "t"), ("gauss", "g")):
        loss = orc.spr_loss(x, y, eps=eps, kind=kind, a=a, b=b, **kw)
        assert abs(loss - float(GOLD[f"loss_{key}{vi}"])) <= 1e-11 * abs(loss)
        nll, mean, var = orc.spr_test_nll(x, y, xt, yt, ym, ys, eps=eps, kind=kind, a=a, b=b, return_parts=True, **kw)
        assert np.abs(mean - GOLD[f"mean{vi}"]).max() <= 1e-9 * np.abs(GOLD[f"mean{vi}"]).max()
        assert np.abs(var - GOLD[f"var{vi}"]).max() <= 1e-9 * np.abs(GOLD[f"var{vi}"]).max()
        assert abs(nll - float(GOLD[f"nll_{key}{vi}"])) <= 1e-9 * abs(nll)


@pytest.mark.parametrize("vi,kw", list(_variants()))
def test_oracle_gradient_matches_mpmath_golden(vi, kw):
    """d SPR.loss / d (w_std, b_std, last_w_std, eps, alpha, beta) against 50-digit mp differentiation."""
    x, y = GOLD["x"], GOLD["y"]
    eps, a, b = float(GOLD["eps"]), float(GOLD["a"]), float(GOLD["b"])
    for kind, key in (("student_t", "t"), ("gauss", "g")):
        loss, grad = orc.spr_loss_grad(x, y, eps=eps, kind=kind, a=a, b=b, **kw)
        assert abs(loss - float(GOLD[f"loss_{key}{vi}"])) <= 1e-11 * abs(loss)
        ref = GOLD[f"grad_{key}{vi}"]
        # eps = 1e-3 on a matrix with a collinear pair: cond ~ 1e4, so ~1e-11 relative on the explicit inverse
        assert np.abs(grad - ref).max() <= 1e-9 * np.abs(ref).max(), (grad, ref)


def test_oracle_gram_dual_matches_finite_differences():
    rng = np.random.default_rng(5)
    x, x2 = rng.standard_normal((40, 6)), rng.standard_normal((17, 6))
    for _, kw in _variants():
        K, dw, db, dv = orc.nngp_gram_dual(x, x2, **kw)
        assert np.abs(K - orc.nngp_gram(x, x2, **kw)).max() <= 1e-14 * np.abs(K).max()
        for name, d in (("w_std", dw), ("b_std", db), ("last_w_std", dv)):
            h = 1e-6
            p, m = dict(kw), dict(kw)
            p[name] += h
            m[name] -= h
            fd = (orc.nngp_gram(x, x2, **p) - orc.nngp_gram(x, x2, **m)) / (2 * h)
            assert np.abs(fd - d).max() <= 1e-8 * max(np.abs(d).max(), 1.0), name


def test_oracle_dual_homogeneity_identities():
    """Analytic identities of the recursion's derivatives: with b_std = 0 the ReLU NNGP kernel is homogeneous of
    degree 2L in w_std (w dK/dw = 2L K, mlp) and every stack is homogeneous of degree 2 in last_w_std."""
    rng = np.random.default_rng(11)
    x, x2 = rng.standard_normal((30, 5)), rng.standard_normal((12, 5))
    for L in (1, 3):
        K, dw, db, dv = orc.nngp_gram_dual(x, x2, num_hiddens=L, act="relu", arch="mlp", w_std=1.3, b_std=0.0,
                                           last_w_std=0.7)
        assert np.abs(1.3 * dw - 2 * L * K).max() <= 1e-13 * np.abs(K).max()
        assert np.abs(db).max() == 0.0                      # d/db of b^2 at b = 0
        assert np.abs(0.7 * dv - 2 * K).max() <= 1e-14 * np.abs(K).max()
    for act, arch in (("erf", "mlp"), ("relu", "resnet"), ("erf", "resnet")):
        K, dw, db, dv = orc.nngp_gram_dual(x, x2, num_hiddens=2, act=act, arch=arch, w_std=0.9, b_std=0.4,
                                           last_w_std=1.6)
        assert np.abs(1.6 * dv - 2 * K).max() <= 1e-14 * np.abs(K).max()


def test_oracle_loss_invariances():
    """SPR.loss does not depend on the order of the data; for the Gaussian likelihood scaling (last_w_std, eps, y)
    by (c, c^2, c) shifts log p by -N log c exactly."""
    x, y, *_ = __import__("tests.synth", fromlist=["regression_data"]).regression_data(120, 7)
    kw = dict(num_hiddens=3, act="relu", arch="mlp", w_std=1.1, b_std=0.3)
    base = orc.spr_loss(x, y, last_w_std=0.8, eps=1e-3, kind="student_t", a=2.5, b=1.5, **kw)
    perm = np.random.default_rng(3).permutation(120)
    assert abs(orc.spr_loss(x[perm], y[perm], last_w_std=0.8, eps=1e-3, kind="student_t", a=2.5, b=1.5, **kw) - base) \
        <= 1e-12 * abs(base)
    c = 1.7
    g0 = orc.spr_loss(x, y, last_w_std=0.8, eps=1e-3, kind="gauss", **kw)
    g1 = orc.spr_loss(x, c * y, last_w_std=0.8 * c, eps=1e-3 * c * c, kind="gauss", **kw)
    assert abs((g1 - g0) - np.log(c)) <= 1e-12
    # gradient consistency with that scaling law: d/dc [loss(c)] at c = 1 = 1 = v dl/dv + 2 eps dl/deps + y.dl/dy; the
    # y-part is -(1/N) y^T d log p / d y = quad / N for the Gaussian
    loss, grad = orc.spr_loss_grad(x, y, last_w_std=0.8, eps=1e-3, kind="gauss", **kw)
    n = 120
    K = orc.nngp_gram(x, last_w_std=0.8, **kw) + 1e-3 * np.eye(n)
    quad = float(y @ np.linalg.solve(K, y))
    assert abs(0.8 * grad[2] + 2 * 1e-3 * grad[3] + quad / n - 1.0) <= 1e-9


def test_c_recursion_matches_numpy():
    rng = np.random.default_rng(0)
    x, x2 = rng.standard_normal((300, 7)), rng.standard_normal((111, 7))
    for _, kw in _variants():
        a = orc.nngp_gram(x, x2, **kw)
        b = orc.nngp_gram(x, x2, fast=True, **kw)
        assert np.abs(a - b).max() <= 5e-15 * np.abs(a).max()


def test_mvt_logpdf_vs_scipy():
    rng = np.random.default_rng(1)
    n = 60
    x = rng.standard_normal((n, 6))
    y = rng.standard_normal(n)
    K = orc.nngp_gram(x, num_hiddens=3, b_std=0.1) + 1e-6 * np.eye(n)
    a, b = 2.0, 3.0
    ours = orc.prior_logpdf(y, K, kind="student_t", a=a, b=b)
    ref = scipy.stats.multivariate_t(loc=np.zeros(n), shape=(b / a) * K, df=2 * a).logpdf(y)
    assert abs(ours - ref) <= 1e-11 * abs(ref)
    ours_g = orc.prior_logpdf(y, K, kind="gauss")
    ref_g = scipy.stats.multivariate_normal(mean=np.zeros(n), cov=K).logpdf(y)
    assert abs(ours_g - ref_g) <= 1e-9 * abs(ref_g)


def test_t_logpdf_vs_scipy():
    x = np.linspace(-3, 3, 11)
    got = orc._t_logpdf(x, 7.5, 0.3, 1.7)
    ref = scipy.stats.t.logpdf(x, 7.5, loc=0.3, scale=1.7)
    assert np.abs(got - ref).max() <= 1e-13


def test_analytic_identities():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((40, 9))
    q = np.einsum("ij,ij->i", x, x) / 9
    # ReLU, sigma_b = 0: diag K = sigma_v^2 sigma_w^(2L) q / 2^L
    L, w, v = 4, 1.3, 0.8
    K = orc.nngp_gram(x, num_hiddens=L, act="relu", w_std=w, b_std=0.0, last_w_std=v)
    assert np.abs(np.diag(K) - v * v * w ** (2 * L) * q / 2 ** L).max() <= 1e-14 * np.abs(K).max()
    assert np.abs(orc.nngp_diag(x, num_hiddens=L, act="relu", w_std=w, b_std=0.0, last_w_std=v) - np.diag(K)).max() <= 1e-14
    # orthogonal inputs, one layer: K = sigma_w^2 sqrt(q1 q2) / (2 pi)
    e = np.zeros((2, 4)); e[0, 0] = 2.0; e[1, 1] = 3.0
    K1 = orc.nngp_gram(e, num_hiddens=1, act="relu", w_std=w, b_std=0.0, last_w_std=1.0)
    assert abs(K1[0, 1] - w * w * math.sqrt((4 / 4) * (9 / 4)) / (2 * math.pi)) <= 1e-15
    # erf diagonal closed form
    Ke = orc.nngp_gram(x, num_hiddens=1, act="erf", w_std=1.0, b_std=0.0, last_w_std=1.0)
    assert np.abs(np.diag(Ke) - (2 / math.pi) * np.arcsin(2 * q / (1 + 2 * q))).max() <= 1e-15
    # symmetry, PSD
    assert np.array_equal(K, K.T) or np.abs(K - K.T).max() <= 1e-16
    assert np.linalg.eigvalsh(K).min() > -1e-10


@pytest.mark.parametrize("act", ["relu", "erf"])
def test_finite_width_monte_carlo(act):
    """The recursion is the infinite-width limit of an NTK-parameterised MLP (what stax.Dense/Relu/Erf mean)."""
    from scipy.special import erf
    rng = np.random.default_rng(3)
    d, width, w_std, b_std = 6, 200_000, 1.2, 0.4
    x = rng.standard_normal((3, d))
    W = rng.standard_normal((d, width))
    bias = rng.standard_normal(width)
    pre = w_std / math.sqrt(d) * (x @ W) + b_std * bias
    h = np.maximum(pre, 0) if act == "relu" else erf(pre)
    mc = h @ h.T / width                       # last Dense with W_std = 1, no bias
    K = orc.nngp_gram(x, num_hiddens=1, act=act, w_std=w_std, b_std=b_std, last_w_std=1.0)
    assert np.abs(mc - K).max() <= 6e-3 * np.abs(K).max()


def test_predict_matches_textbook_gp():
    rng = np.random.default_rng(4)
    x, xt, y = rng.standard_normal((50, 5)), rng.standard_normal((7, 5)), rng.standard_normal(50)
    kw = dict(num_hiddens=2, act="relu", w_std=1.0, b_std=0.2, last_w_std=1.0)
    mean, cov = orc.nt_predict(x, y, xt, 1e-3, kernel_kwargs=kw)
    K = orc.nngp_gram(x, **kw)
    A = K + 1e-3 * np.trace(K) / 50 * np.eye(50)
    Ktd = orc.nngp_gram(xt, x, **kw)
    assert np.allclose(mean[:, 0], Ktd @ np.linalg.solve(A, y), rtol=1e-10, atol=1e-12)
    assert np.allclose(cov, orc.nngp_gram(xt, **kw) - Ktd @ np.linalg.solve(A, Ktd.T), rtol=1e-9, atol=1e-12)


def test_non_pd_is_nan_not_an_exception():
    K = np.array([[1.0, 2.0], [2.0, 1.0]])
    assert math.isnan(orc.prior_logpdf(np.ones(2), K, kind="student_t", a=2.0, b=2.0))
    assert math.isnan(orc.prior_logpdf(np.ones(2), K, kind="gauss"))


def test_softplus_roundtrip():
    v = np.array([1e-8, 1e-6, 0.3, 1.0, 2.0, 25.0])
    assert np.allclose(orc.softplus(orc.softplus_inverse(v)), v, rtol=1e-12, atol=0)


def test_unsupported_act_and_arch():
    x = np.ones((2, 2))
    with pytest.raises(KeyError):
        orc.nngp_gram(x, num_hiddens=1, act="tanh")
    with pytest.raises(ValueError):
        orc.nngp_gram(x, num_hiddens=1, arch="cnn")


def test_blocked_cholesky_path_matches_one_call_lapack(monkeypatch):
    """above N = 40 000 the oracle factors block by block (SciPy's 32-bit LAPACK interface segfaults past 2^31 elements);
    forced on at a small size it must reproduce the one-call path: same loss, same factor"""
    import scipy.linalg as sla
    from tests.synth import regression_data
    x, y, *_ = regression_data(700, 6)
    kw = dict(num_hiddens=2, act="relu", arch="mlp", w_std=1.1, b_std=0.2, last_w_std=0.9, eps=1e-4, a=2.0, b=3.0)
    ref_t = orc.spr_loss(x, y, kind="student_t", **kw)
    ref_g = orc.spr_loss(x, y, kind="gauss", **kw)
    monkeypatch.setattr(orc, "BLOCKED_ABOVE", 100)
    assert abs(orc.spr_loss(x, y, kind="student_t", **kw) - ref_t) <= 1e-13 * abs(ref_t)
    assert abs(orc.spr_loss(x, y, kind="gauss", **kw) - ref_g) <= 1e-13 * abs(ref_g)
    k = orc.nngp_gram(x, num_hiddens=2, act="relu", w_std=1.1, b_std=0.2, last_w_std=0.9) + 1e-4 * np.eye(700)
    L = orc.cholesky_blocked_inplace(k.copy(), nb=128)
    assert np.abs(np.tril(L) - sla.cholesky(k, lower=True)).max() <= 1e-12
    z = orc.forward_substitution_blocked(L, y, nb=128)
    assert np.abs(z - sla.solve_triangular(np.tril(L), y, lower=True)).max() <= 1e-10 * np.abs(z).max()


# ---- host-side pieces of the product that need no GPU ---------------------------------------------------
def test_cabi_library_loads_and_exports_every_declared_symbol():
    import smnngp_b200 as sm
    path = sm._lib.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "smnngp.h")).read()
    declared = set(re.findall(r"\b(smnngp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/smnngp.h but not exported"
    assert set(sm._lib.EXPORTS) <= declared
    assert lib.smnngp_abi_version() == 1
    lib.smnngp_lml_workspace_bytes.restype = ctypes.c_size_t
    lib.smnngp_lml_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    nbytes = lib.smnngp_lml_workspace_bytes(60000, 784, 3, 0)
    # factorisation buffer + the panel buffer / block inverse of the single-launch panel solve (512 doubles per row)
    assert 60001 * 60000 * 8 <= nbytes <= 60001 * (60016 + 512) * 8 + (1 << 23)


def test_product_path_fails_loudly_without_gpu():
    import torch
    import smnngp_b200 as sm
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        sm.device.lml(np.zeros((4, 2)), np.zeros(4), spec=sm.StackSpec(1), hp=np.ones(6))


def test_spax_mirror_scalars_and_factories():
    import smnngp_b200 as sm
    from smnngp_b200.spax import NNGPKernel, StudentTLikelihood, ConstraintTrainVar, positive
    v = ConstraintTrainVar(1e-6, constraint=positive())
    assert abs(v.safe_value - 1e-6) <= 1e-18
    assert abs(float(v.value) - float(orc.softplus_inverse(1e-6))) <= 1e-12
    kern = NNGPKernel(lambda w, b, l: sm.get_mlp_kernel(3, act="relu", w_std=w, b_std=b, last_w_std=l), 1.0, 1e-8, 1.0)
    w, b, l = kern.get_params()
    assert abs(w - 1.0) < 1e-12 and abs(b - 1e-8) < 1e-20 and abs(l - 1.0) < 1e-12
    fn = kern.get_kernel_fn()
    assert fn.spec == sm.StackSpec(3, "relu", "mlp") and abs(fn.b_std - 1e-8) < 1e-20
    lik = StudentTLikelihood(2.0, 2.0)
    assert lik.require == ["cov_data", "y_data"] and abs(lik.a.safe_value - 2.0) < 1e-12
    with pytest.raises(KeyError):
        sm.get_mlp_kernel(3, act="tanh")
    assert sm.get_dense_resnet_kernel(2, act="erf").spec.arch == "resnet"
