"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances are BASELINE.json's: Gram entries 1e-10 (|dK| <= 1e-10 * max|K|), LML and predictive moments 1e-8
relative."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import nngp_oracle as orc
from tests.synth import regression_data, pixel_data, DEFAULT_HP

pytestmark = pytest.mark.gpu

GRAM_TOL = 1e-10
LML_TOL = 1e-8


@pytest.fixture(scope="module")
def sm():
    import torch
    import smnngp_b200 as s
    assert torch.cuda.is_available()
    s._lib.load()
    return s


def _hp(sm, **over):
    hp = dict(DEFAULT_HP)
    hp.update(over)
    return hp, sm.make_hp(hp["w_std"], hp["b_std"], hp["last_w_std"], hp["eps"], hp["alpha"], hp["beta"])


def _kw(hp, L, act, arch):
    return dict(num_hiddens=L, act=act, w_std=hp["w_std"], b_std=hp["b_std"], last_w_std=hp["last_w_std"], arch=arch)


@pytest.mark.parametrize("n,d", [(300, 13), (257, 8), (130, 784), (1, 5), (127, 3), (640, 64)])
@pytest.mark.parametrize("act,arch,L,b_std", [("relu", "mlp", 3, 1e-8), ("erf", "mlp", 3, 0.3),
                                               ("relu", "resnet", 2, 0.3), ("erf", "resnet", 1, 0.3),
                                               ("relu", "mlp", 1, 0.0), ("relu", "mlp", 10, 0.1)])
def test_gram_symmetric(sm, n, d, act, arch, L, b_std):
    import torch
    x, *_ = regression_data(n, d, seed=3)
    hp, hpd = _hp(sm, b_std=b_std, w_std=1.3, last_w_std=0.7)
    ref = orc.nngp_gram(x, **_kw(hp, L, act, arch))
    xd = torch.from_numpy(x).cuda()
    k = sm.device.gram(xd, spec=sm.StackSpec(L, act, arch), hp=hpd).cpu().numpy()
    err = np.abs(k - ref).max() / np.abs(ref).max()
    assert err <= GRAM_TOL, f"gram err {err:.3e}"
    assert np.array_equal(k, k.T), "mirrored output must be exactly symmetric"
    # lower-only output with the absolute shift on the diagonal
    kl = sm.device.gram(xd, spec=sm.StackSpec(L, act, arch), hp=hpd, shift="eps_abs", lower_only=True).cpu().numpy()
    assert np.all(np.triu(kl, 1) == 0.0)
    ref_l = np.tril(ref + hp["eps"] * np.eye(n))
    assert np.abs(kl - ref_l).max() / np.abs(ref).max() <= GRAM_TOL
    # diagonal helper
    q = sm.device.nngp_diag(xd, spec=sm.StackSpec(L, act, arch), hp=hpd).cpu().numpy()
    qref = orc.nngp_diag(x, **_kw(hp, L, act, arch))
    assert np.abs(q - qref).max() <= 1e-12 * np.abs(qref).max()


@pytest.mark.parametrize("n,m,d", [(200, 333, 13), (129, 64, 8), (50, 700, 784)])
@pytest.mark.parametrize("act,arch", [("relu", "mlp"), ("erf", "resnet")])
def test_gram_cross(sm, n, m, d, act, arch):
    import torch
    x, _, x2, *_ = regression_data(n, d, t=m, seed=4)
    hp, hpd = _hp(sm, b_std=0.2)
    ref = orc.nngp_gram(x, x2, **_kw(hp, 3, act, arch))
    k = sm.device.gram(torch.from_numpy(x).cuda(), torch.from_numpy(x2).cuda(), spec=sm.StackSpec(3, act, arch),
                       hp=hpd).cpu().numpy()
    assert k.shape == (n, m)
    assert np.abs(k - ref).max() / np.abs(ref).max() <= GRAM_TOL


@pytest.mark.parametrize("sr", [1, 2, 3])
def test_gram_super_tile_walk(sm, sr):
    """L2-aware rasterisation of the Gram kernel (super-tiles of sr x 2 sr tiles) forced on at small sizes: ragged
    edges, more super-tile slots than tiles, symmetric (block-triangular list) and cross (rectangular) cases"""
    import torch
    lib = sm._lib.load()
    lib.smnngp_set_gram_super_rows(sr, 0)
    try:
        hp, hpd = _hp(sm, b_std=0.2)
        spec = sm.StackSpec(2, "relu", "mlp")
        kw = _kw(hp, 2, "relu", "mlp")
        for n in (1000, 1153, 130):
            x, *_ = regression_data(n, 16, seed=n)
            ref = orc.nngp_gram(x, **kw)
            xd = torch.from_numpy(x).cuda()
            k = sm.device.gram(xd, spec=spec, hp=hpd).cpu().numpy()
            assert np.abs(k - ref).max() <= GRAM_TOL * np.abs(ref).max()
            kl = sm.device.gram(xd, spec=spec, hp=hpd, shift="eps_abs", lower_only=True).cpu().numpy()
            assert np.all(np.triu(kl, 1) == 0.0)
            assert np.abs(kl - np.tril(ref + hp["eps"] * np.eye(n))).max() <= GRAM_TOL * np.abs(ref).max()
        x1, *_ = regression_data(700, 16, seed=1)
        x2, *_ = regression_data(1333, 16, seed=2)
        ref = orc.nngp_gram(x1, x2, **kw)
        k = sm.device.gram(torch.from_numpy(x1).cuda(), torch.from_numpy(x2).cuda(), spec=spec, hp=hpd).cpu().numpy()
        assert np.abs(k - ref).max() <= GRAM_TOL * np.abs(ref).max()
    finally:
        lib.smnngp_set_gram_super_rows(8, -1)


def test_gram_duplicate_and_zero_rows(sm):
    """edge cases of the arc-cosine step: identical rows (s -> 0) and all-zero rows (atan2(0, 0) -> pi/2 fill)."""
    import torch
    rng = np.random.default_rng(0)
    x = rng.standard_normal((70, 6))
    x[5] = x[9]
    x[11] = 0.0
    x[12] = 0.0
    hp, hpd = _hp(sm, b_std=0.0)
    for act in ("relu", "erf"):
        ref = orc.nngp_gram(x, **_kw(hp, 3, act, "mlp"))
        k = sm.device.gram(torch.from_numpy(x).cuda(), spec=sm.StackSpec(3, act, "mlp"), hp=hpd).cpu().numpy()
        assert np.isfinite(k).all()
        assert np.abs(k - ref).max() <= GRAM_TOL * np.abs(ref).max()


@pytest.mark.parametrize("n,extra", [(1, 0), (100, 0), (128, 1), (129, 3), (300, 0), (1000, 17), (2500, 1),
                                     # outer panels of 256 / 512 columns: look-ahead + single-launch panel solve with the
                                     # assembled block inverse (2 / 4 diagonal blocks), ragged last panel, carried rows
                                     (4000, 0), (4100, 3), (8704, 2), (9000, 5)])
def test_potrf_vs_lapack(sm, n, extra):
    import torch
    rng = np.random.default_rng(n)
    b = rng.standard_normal((n, n + 8))
    a = b @ b.T / (n + 8) + 1e-3 * np.eye(n)
    r = rng.standard_normal((extra, n))
    buf = np.vstack([a, r])
    ad = torch.from_numpy(buf).cuda()
    logdet, info = sm.device.potrf_(ad, n)
    L = sla.cholesky(a, lower=True)
    got = ad.cpu().numpy()
    assert int(info.item()) == 0
    errL = np.abs(np.tril(got[:n]) - L).max() / np.abs(L).max()
    assert errL <= 1e-11, f"L err {errL:.3e}"
    # strict upper triangle untouched
    assert np.array_equal(np.triu(got[:n], 1), np.triu(a, 1))
    assert abs(logdet.item() - np.log(np.diag(L)).sum()) <= 1e-10 * max(1.0, abs(np.log(np.diag(L)).sum()))
    if extra:
        want = sla.solve_triangular(L, r.T, lower=True).T
        errR = np.abs(got[n:] - want).max() / np.abs(want).max()
        assert errR <= 1e-9, f"carried rows err {errR:.3e}"


def test_potrf_not_positive_definite(sm):
    import torch
    rng = np.random.default_rng(1)
    n = 300
    b = rng.standard_normal((n, n))
    a = b @ b.T / n
    a[200, 200] = -1.0
    ad = torch.from_numpy(a.copy()).cuda()
    logdet, info = sm.device.potrf_(ad)
    assert int(info.item()) == 201          # LAPACK convention: 1 + first bad pivot
    assert np.isnan(ad.cpu().numpy()[250, 250])


@pytest.mark.parametrize("fused", [1, 0])
def test_potrf_fused_panel_matches_block_substitution(sm, fused):
    """the single-launch panel solve (full inverse of the diagonal block, out of place) and the round-1 in-place
    128-block substitution are two evaluations of the same factorisation: both must match LAPACK, also when the
    matrix is not positive definite in a late panel (NaN + info, nothing raises)"""
    import torch
    n = 5000
    rng = np.random.default_rng(3)
    b = rng.standard_normal((n, n + 8))
    a = b @ b.T / (n + 8) + 1e-3 * np.eye(n)
    lib = sm._lib.load()
    lib.smnngp_set_fused_panel(fused)
    try:
        ad = torch.from_numpy(a.copy()).cuda()
        logdet, info = sm.device.potrf_(ad)
        L = sla.cholesky(a, lower=True)
        assert int(info.item()) == 0
        assert np.abs(np.tril(ad.cpu().numpy()) - L).max() / np.abs(L).max() <= 1e-11
        a[4500, 4500] = -1.0
        ad = torch.from_numpy(a.copy()).cuda()
        logdet, info = sm.device.potrf_(ad)
        assert int(info.item()) == 4501 and np.isnan(ad.cpu().numpy()[4900, 4800])
    finally:
        lib.smnngp_set_fused_panel(1)


@pytest.mark.parametrize("n,d,L,act,arch,kind", [
    (506, 13, 3, "relu", "mlp", "student_t"),     # BASELINE config 1
    (506, 13, 3, "relu", "mlp", "gauss"),
    (404, 13, 4, "erf", "mlp", "student_t"),
    (1300, 8, 3, "relu", "resnet", "student_t"),
    (2100, 8, 3, "relu", "mlp", "student_t"),     # panel width 256 path
    (77, 5, 2, "relu", "mlp", "gauss"),
])
def test_lml_parity(sm, n, d, L, act, arch, kind):
    import torch
    x, y, *_ = regression_data(n, d)
    b_std = 0.3 if act == "erf" else DEFAULT_HP["b_std"]
    hp, hpd = _hp(sm, b_std=b_std)
    ref_loss = orc.spr_loss(x, y, kind=kind, a=hp["alpha"], b=hp["beta"], eps=hp["eps"], **_kw(hp, L, act, arch))
    out, info = sm.device.lml(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), spec=sm.StackSpec(L, act, arch),
                              hp=hpd, kind=kind)
    out = out.cpu().numpy()
    assert int(info.item()) == 0
    assert abs(out[1] - ref_loss) <= LML_TOL * abs(ref_loss), f"loss {out[1]} vs {ref_loss}"
    assert abs(out[0] + ref_loss * n) <= LML_TOL * abs(ref_loss * n)
    # host-buffer C-ABI entry point
    hph = np.array([hp["w_std"], hp["b_std"], hp["last_w_std"], hp["eps"], hp["alpha"], hp["beta"]])
    out_h, info_h = sm.device.lml(x, y, spec=sm.StackSpec(L, act, arch), hp=hph, kind=kind)
    assert info_h == 0 and out_h[1] == out[1], "host and device entry points must agree bitwise"


def test_lml_non_pd_gives_nan(sm):
    """duplicate rows + eps -> 0 makes K singular: reference yields NaN (jax cholesky), never raises."""
    import torch
    x, y, *_ = regression_data(200, 4)
    x[10] = x[20]
    x[30] = x[20]
    hp, hpd = _hp(sm, eps=1e-300)
    out, info = sm.device.lml(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), spec=sm.StackSpec(3, "relu", "mlp"),
                              hp=hpd)
    if int(info.item()) != 0:
        assert np.isnan(out.cpu().numpy()[0])


@pytest.mark.parametrize("n,t,d,c,act,arch", [(404, 52, 13, 1, "relu", "mlp"), (900, 130, 8, 3, "erf", "mlp"),
                                              (1500, 257, 8, 1, "relu", "resnet")])
def test_predict_parity(sm, n, t, d, c, act, arch):
    import torch
    x, y, xt, yt, *_ = regression_data(n, d, t=t)
    rng = np.random.default_rng(5)
    Y = y[:, None] if c == 1 else np.column_stack([y] + [rng.standard_normal(n) for _ in range(c - 1)])
    hp, hpd = _hp(sm, b_std=0.3 if act == "erf" else 1e-8)
    kw = _kw(hp, 3, act, arch)
    mean_ref, cov_ref = orc.nt_predict(x, Y, xt, hp["eps"], kernel_kwargs=kw)
    mean, var, info = sm.device.predict(torch.from_numpy(x).cuda(), torch.from_numpy(Y).cuda(),
                                        torch.from_numpy(xt).cuda(), spec=sm.StackSpec(3, act, arch), hp=hpd)
    assert int(info.item()) == 0
    mean, var = mean.cpu().numpy(), var.cpu().numpy()
    assert np.abs(mean - mean_ref).max() <= LML_TOL * np.abs(mean_ref).max()
    vref = np.diag(cov_ref)
    # variance = k_tt - ||v||^2 cancels: relative 1e-8 on the variance with an absolute floor tied to k_tt
    ktt = orc.nngp_diag(xt, **kw)
    assert np.all(np.abs(var - vref) <= LML_TOL * np.abs(vref) + 1e-13 * ktt)
    # full covariance, as neural_tangents returns it
    mean_c, cov, _ = sm.device.predict(torch.from_numpy(x).cuda(), torch.from_numpy(Y).cuda(), torch.from_numpy(xt).cuda(),
                                       spec=sm.StackSpec(3, act, arch), hp=hpd, full_cov=True)
    cov = cov.cpu().numpy()
    assert np.abs(cov - cov_ref).max() <= LML_TOL * np.abs(cov_ref).max() + 1e-13 * ktt.max()
    assert np.array_equal(mean_c.cpu().numpy(), mean)
    # host entry point
    hph = np.array([hp[k] for k in ("w_std", "b_std", "last_w_std", "eps", "alpha", "beta")])
    mean_h, var_h, info_h = sm.device.predict(x, Y, xt, spec=sm.StackSpec(3, act, arch), hp=hph)
    assert info_h == 0 and np.array_equal(mean_h, mean) and np.array_equal(var_h, var)


@pytest.mark.parametrize("n,t,d,kind", [(404, 52, 13, "student_t"), (404, 52, 13, "gauss"), (1100, 300, 8, "student_t")])
def test_test_nll_parity(sm, n, t, d, kind):
    import torch
    x, y, xt, yt, ym, ys = regression_data(n, d, t=t)
    hp, hpd = _hp(sm)
    kw = _kw(hp, 3, "relu", "mlp")
    ref, mref, vref = orc.spr_test_nll(x, y, xt, yt, ym, ys, eps=hp["eps"], kind=kind, a=hp["alpha"], b=hp["beta"],
                                       return_parts=True, **kw)
    nll, mean, var, info = sm.device.test_nll(*(torch.from_numpy(v).cuda() for v in (x, y, xt, yt)), ym, ys,
                                              spec=sm.StackSpec(3, "relu", "mlp"), hp=hpd, kind=kind)
    assert int(info.item()) == 0
    assert abs(nll.item() - ref) <= LML_TOL * abs(ref), f"{nll.item()} vs {ref}"
    hph = np.array([hp[k] for k in ("w_std", "b_std", "last_w_std", "eps", "alpha", "beta")])
    nll_h, *_ = sm.device.test_nll(x, y, xt, yt, ym, ys, spec=sm.StackSpec(3, "relu", "mlp"), hp=hph, kind=kind)
    assert nll_h == nll.item()


def test_spax_api_dropin(sm):
    """The reference's call pattern (regression/train.py:126-142, spax/models.py) on the mirrored API."""
    import torch
    from smnngp_b200.spax import NNGPKernel, StudentTLikelihood, GaussianLikelihood, SPR
    x, y, xt, yt, ym, ys = regression_data(404, 13, t=52)
    base = sm.get_mlp_kernel

    def get_kernel_fn(w_std, b_std, last_w_std):
        return base(3, act="relu", w_std=w_std, b_std=b_std, last_w_std=last_w_std)

    xd, yd, xtd, ytd = (torch.from_numpy(v).cuda() for v in (x, y, xt, yt))
    for lik, kind in ((StudentTLikelihood(2.0, 2.0), "student_t"), (GaussianLikelihood(), "gauss")):
        kernel = NNGPKernel(get_kernel_fn, 1.0, 1e-8, 1.0)
        model = SPR(kernel, lik, xd, yd, ym, ys, eps=1e-6)
        w, b, v = kernel.get_params()
        kw = dict(num_hiddens=3, act="relu", w_std=w, b_std=b, last_w_std=v, arch="mlp")
        ref_loss = orc.spr_loss(x, y, eps=model.eps.safe_value, kind=kind, a=2.0, b=2.0, **kw)
        ref_nll = orc.spr_test_nll(x, y, xt, yt, ym, ys, eps=model.eps.safe_value, kind=kind, a=2.0, b=2.0, **kw)
        loss = float(model.loss())
        nll = float(model.test_nll(xtd, ytd))
        assert abs(loss - ref_loss) <= LML_TOL * abs(ref_loss)
        assert abs(nll - ref_nll) <= LML_TOL * abs(ref_nll)
        # un-fused composition (K + jitter -> prior_logpdf; predict -> logpdf), as the reference writes it
        kernel_fn = kernel.get_kernel_fn()
        cov = kernel.K(kernel_fn, xd) + sm.spax.jitter(404, eps=model.eps.safe_value)
        lp = float(lik.prior_logpdf(yd, cov))
        assert abs(-lp / 404 - ref_loss) <= LML_TOL * abs(ref_loss)
        mean, var = kernel.predict(kernel_fn, xd, yd[:, None], xtd, eps=model.eps.safe_value)
        aux = (kernel.K(kernel_fn, xd), yd) if lik.require else None
        logp = lik.logpdf(ytd * ys + ym, mean.flatten() * ys + ym, var * ys ** 2, aux)
        assert abs(-float(logp.mean()) - ref_nll) <= LML_TOL * abs(ref_nll)
    # numpy inputs go through the host-buffer C-ABI
    model_h = SPR(NNGPKernel(get_kernel_fn, 1.0, 1e-8, 1.0), StudentTLikelihood(2.0, 2.0), x, y, ym, ys)
    assert isinstance(model_h.loss(), float)


def test_medium_lml_pixel_shape(sm):
    """MNIST-shaped inputs at a size the oracle still finishes in seconds (D = 784, two outer panels)."""
    import torch
    x, y, *_ = pixel_data(3000, 784)
    hp, hpd = _hp(sm)
    ref = orc.spr_loss(x, y, eps=hp["eps"], kind="student_t", a=2.0, b=2.0, **_kw(hp, 3, "relu", "mlp"))
    out, info = sm.device.lml(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), spec=sm.StackSpec(3, "relu", "mlp"),
                              hp=hpd)
    assert int(info.item()) == 0
    assert abs(out[1].item() - ref) <= LML_TOL * abs(ref)


def test_golden_vectors_on_gpu(sm):
    """The committed mpmath fixtures (tests/golden/make_golden.py), CUDA path directly against them."""
    import os
    import torch
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nngp_golden.npz"))
    x, xt, y, yt = (torch.from_numpy(g[k]).cuda() for k in ("x", "xt", "y", "yt"))
    eps, a, b = float(g["eps"]), float(g["a"]), float(g["b"])
    for vi, s in enumerate(g["variants"]):
        act, arch, L, w, bs, v = str(s).split(",")
        spec = sm.StackSpec(int(L), act, arch)
        hpd = sm.make_hp(float(w), float(bs), float(v), eps, a, b)
        K = sm.device.gram(x, spec=spec, hp=hpd).cpu().numpy()
        assert np.abs(K - g[f"K{vi}"]).max() <= GRAM_TOL * np.abs(K).max()
        Ktd = sm.device.gram(xt, x, spec=spec, hp=hpd).cpu().numpy()
        assert np.abs(Ktd - g[f"Ktd{vi}"]).max() <= GRAM_TOL * np.abs(K).max()
        for kind, key in (("student_t", "t"), ("gauss", "g")):
            out, info = sm.device.lml(x, y, spec=spec, hp=hpd, kind=kind)
            assert abs(out[1].item() - float(g[f"loss_{key}{vi}"])) <= LML_TOL * abs(float(g[f"loss_{key}{vi}"]))
            nll, mean, var, info = sm.device.test_nll(x, y, xt, yt, float(g["y_mean"]), float(g["y_std"]), spec=spec,
                                                      hp=hpd, kind=kind)
            assert np.abs(mean.cpu().numpy() - g[f"mean{vi}"]).max() <= LML_TOL * np.abs(g[f"mean{vi}"]).max()
            assert np.abs(var.cpu().numpy() - g[f"var{vi}"]).max() <= LML_TOL * np.abs(g[f"var{vi}"]).max()
            assert abs(nll.item() - float(g[f"nll_{key}{vi}"])) <= LML_TOL * abs(float(g[f"nll_{key}{vi}"]))


@pytest.mark.parametrize("m,n,k,lower", [(2745, 2744, 256, 1), (2816, 2752, 256, 0), (4096, 4160, 16, 0),
                                         (4096, 4096, 128, 0), (8192, 8256, 512, 1), (1000, 130, 48, 0),
                                         # short K, many tiles per CTA: the release of a tile's LAST slab is followed by
                                         # the epilogue, not by another slab (the window the delayed release left open)
                                         (12288, 12288, 16, 0), (12288, 12288, 32, 1), (16384, 8192, 48, 0),
                                         (9000, 9100, 128, 1)])
def test_update_kernel_many_tiles_per_cta(sm, m, n, k, lower):
    """C -= A B^T on the TMA-fed persistent kernel when every math group walks through several tiles (regression
    test for the stage-release race: a stage released while ld.shared was in flight got overwritten by TMA)."""
    import torch
    from smnngp_b200.distributed import CudaBackend
    be = CudaBackend("cuda")
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, dtype=torch.float64, device="cuda")
    b = torch.randn(n, k, dtype=torch.float64, device="cuda")
    c0 = torch.randn(m, n, dtype=torch.float64, device="cuda")
    ref = c0 - a @ b.T
    if lower:
        mask = torch.arange(n, device="cuda")[None, :] <= torch.arange(m, device="cuda")[:, None]
        ref = torch.where(mask, ref, c0)
    for _ in range(4):
        c = c0.clone()
        be.update(a, b, c, bool(lower), 0, 1, 0)
        assert float((c - ref).abs().max()) <= 1e-11 * k


def test_lml_cuda_graph_replay(sm):
    """the fused call is enqueue-only: capture once, replay with new inputs / hyper-parameters"""
    import torch
    x, y, *_ = regression_data(1300, 8)
    spec = sm.StackSpec(3, "relu", "mlp")
    g = sm.device.LmlGraph(1300, 8, spec=spec)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    for b_std in (1e-8, 0.3):
        hp, hpd = _hp(sm, b_std=b_std)
        out, info = g(xd, yd, hpd)
        ref = orc.spr_loss(x, y, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], **_kw(hp, 3, "relu", "mlp"))
        assert int(info.item()) == 0 and abs(out[1].item() - ref) <= LML_TOL * abs(ref)
        direct, _ = sm.device.lml(xd, yd, spec=spec, hp=hpd)
        assert out[1].item() == direct[1].item()


@pytest.mark.parametrize("kind", ["student_t", "gauss"])
def test_draw_stage(sm, kind):
    """sample_f_iid (spax/priors.py:30-36, :60-68) + test_log_likelihood / get_correct_count (spax/utils.py:61-74):
    distribution of the draws, and the fused metrics kernel against the oracle evaluated on the SAME draws."""
    import torch
    import scipy.stats
    rng = np.random.default_rng(7)
    T, C, S = 37, 10, 4000
    mean = rng.standard_normal((T, C))
    var = rng.uniform(0.05, 2.0, T)
    label = rng.integers(0, C, T)
    a, b = 2.5, 1.5
    hpd = sm.make_hp(1.0, 0.0, 1.0, 1e-6, a, b)
    md, vd = torch.from_numpy(mean).cuda(), torch.from_numpy(var).cuda()
    f = sm.device.sample_f_iid(md, vd, hp=hpd, kind=kind, num_samples=S, seed=1234).cpu().numpy()   # [C, T, S]
    assert f.shape == (C, T, S) and np.isfinite(f).all()
    scale = np.sqrt((b / a if kind == "student_t" else 1.0) * var)
    z = (f - mean.T[:, :, None]) / scale[None, :, None]
    dist = scipy.stats.t(2 * a) if kind == "student_t" else scipy.stats.norm()
    ks = scipy.stats.kstest(z.ravel()[:200000], dist.cdf)
    assert ks.pvalue > 1e-3, ks
    assert abs(np.corrcoef(z[0, 0], z[1, 0])[0, 1]) < 0.08           # iid across classes
    assert abs(np.corrcoef(z[0, 0], z[0, 1])[0, 1]) < 0.08           # ... and across test points
    nll, correct, ll, pred = sm.device.draw_metrics(md, vd, label, hp=hpd, kind=kind, num_samples=S, seed=1234)
    ref_nll = -orc.test_log_likelihood(f, label)
    ref_correct = orc.get_correct_count(f, label)
    assert abs(nll.item() - ref_nll) <= 1e-10 * abs(ref_nll)
    assert int(correct.item()) == ref_correct
    # the mirrored prior class
    from smnngp_b200.spax import InverseGammaPrior, GaussianPrior
    prior = InverseGammaPrior(a, b) if kind == "student_t" else GaussianPrior()
    f2 = prior.sample_f_iid(1234, md.T.contiguous(), vd, S).cpu().numpy()
    assert np.array_equal(f2, f)


def test_find_grid_point(sm):
    """experiments/regression/find.py:134-160, one (w_std, b_std, eps) point: predictive (relative regulariser) +
    log det / quadratic form (absolute jitter)."""
    import torch
    x, y, xt, yt, *_ = regression_data(404, 13, t=52)
    hp, hpd = _hp(sm, w_std=1.5, b_std=0.2, eps=1e-3)
    kw = _kw(hp, 3, "relu", "mlp")
    mean_ref, var_ref, logdet_ref, quad_ref = orc.find_grid_point(x, y, xt, hp["eps"], kernel_kwargs=kw)
    mean, var, logdet, quad, info = sm.device.grid_point(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                                         torch.from_numpy(xt).cuda(), spec=sm.StackSpec(3, "relu", "mlp"),
                                                         hp=hpd)
    assert int(info.item()) == 0
    assert np.abs(mean.cpu().numpy() - mean_ref).max() <= LML_TOL * np.abs(mean_ref).max()
    assert np.abs(var.cpu().numpy() - var_ref).max() <= LML_TOL * np.abs(var_ref).max()
    assert abs(logdet.item() - logdet_ref) <= LML_TOL * abs(logdet_ref)
    assert abs(quad.item() - quad_ref) <= LML_TOL * abs(quad_ref)


@pytest.mark.parametrize("n,t,d,act,arch,L", [(404, 52, 13, "relu", "mlp", 3), (900, 130, 8, "erf", "mlp", 2),
                                              (1500, 257, 16, "relu", "resnet", 2)])
def test_grid_search_with_cached_base(sm, n, t, d, act, arch, L):
    """find.py grid with the base Gram cached: every point must equal the from-scratch evaluation."""
    import torch
    x, y, xt, yt, *_ = regression_data(n, d, t=t)
    xd, yd, xtd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(xt).cuda()
    spec = sm.StackSpec(L, act, arch)
    gs = sm.device.GridSearch(xd, yd, xtd, spec=spec)
    for w_std, b_std, eps in ((1.0, 0.3, 1e-3), (1.7, 0.05, 1e-2), (0.6, 0.9, 1e-4)):
        hp, hpd = _hp(sm, w_std=w_std, b_std=b_std, eps=eps)
        mean, var, logdet, quad, info = gs.point(hpd)
        mean_ref, var_ref, logdet_ref, quad_ref = orc.find_grid_point(x, y, xt, eps, kernel_kwargs=_kw(hp, L, act, arch))
        assert int(info.item()) == 0
        assert np.abs(mean.cpu().numpy() - mean_ref).max() <= LML_TOL * np.abs(mean_ref).max()
        ktt = orc.nngp_diag(xt, **_kw(hp, L, act, arch))
        assert np.all(np.abs(var.cpu().numpy() - var_ref) <= LML_TOL * np.abs(var_ref) + 1e-13 * ktt)
        assert abs(logdet.item() - logdet_ref) <= LML_TOL * abs(logdet_ref)
        assert abs(quad.item() - quad_ref) <= LML_TOL * abs(quad_ref)
        m2, v2, ld2, q2, _ = sm.device.grid_point(xd, yd, xtd, spec=spec, hp=hpd)      # from scratch
        assert np.abs((mean - m2).cpu().numpy()).max() <= 1e-11 * np.abs(mean_ref).max()
        assert abs(logdet.item() - ld2.item()) <= 1e-11 * abs(logdet_ref)


# ---- BASELINE configs at (or near) their named sizes --------------------------------------------------------------
def test_c2_uci_shape_lml_and_test_nll(sm):
    """BASELINE config 2: N = 10 000, D = 8, T = 1000, 3-layer ReLU, Student-t - SPR.loss and SPR.test_nll against the
    oracle (spax/models.py:93-120).  The oracle needs ~15 s of host time at this size."""
    import torch
    n, d, t = 10000, 8, 1000
    x, y, xt, yt, ym, ys = regression_data(n, d, t=t)
    hp, hpd = _hp(sm)
    spec = sm.StackSpec(3, "relu", "mlp")
    kw = _kw(hp, 3, "relu", "mlp")
    xd, yd, xtd, ytd = (torch.from_numpy(v).cuda() for v in (x, y, xt, yt))
    out, info = sm.device.lml(xd, yd, spec=spec, hp=hpd)
    ref = orc.spr_loss(x, y, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], fast=True, **kw)
    assert int(info.item()) == 0
    assert abs(out[1].item() - ref) <= LML_TOL * abs(ref), (out[1].item(), ref)
    nll, mean, var, info = sm.device.test_nll(xd, yd, xtd, ytd, ym, ys, spec=spec, hp=hpd)
    ref_nll = orc.spr_test_nll(x, y, xt, yt, ym, ys, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], **kw)
    assert int(info.item()) == 0
    assert abs(float(nll.item()) - ref_nll) <= LML_TOL * abs(ref_nll), (float(nll.item()), ref_nll)


def test_c4_cifar_shape_predict(sm):
    """BASELINE config 4 shape (D = 3072, C = 10 one-hot-centred targets) at N = 8192, T = 2048: NNGPKernel.predict
    (spax/kernels.py:29-32) - mean and diag(cov) against the oracle (not against another GPU library)."""
    import torch
    n, d, t, c = 8192, 3072, 2048, 10
    rng = np.random.default_rng(10)
    x = rng.standard_normal((n + t, d))
    lab = rng.integers(0, c, n)
    Y = np.eye(c)[lab] - 1.0 / c
    x, xt = np.ascontiguousarray(x[:n]), np.ascontiguousarray(x[n:])
    hp, hpd = _hp(sm, eps=1e-4)
    spec = sm.StackSpec(3, "relu", "mlp")
    kw = _kw(hp, 3, "relu", "mlp")
    mean_ref, cov_ref = orc.nt_predict(x, Y, xt, hp["eps"], kernel_kwargs=kw)
    vref, ktt = np.diag(cov_ref), orc.nngp_diag(xt, **kw)
    mean, var, info = sm.device.predict(torch.from_numpy(x).cuda(), torch.from_numpy(Y).cuda(),
                                        torch.from_numpy(xt).cuda(), spec=spec, hp=hpd)
    assert int(info.item()) == 0
    mean, var = mean.cpu().numpy(), var.cpu().numpy()
    assert np.abs(mean - mean_ref).max() <= LML_TOL * np.abs(mean_ref).max()
    assert np.all(np.abs(var - vref) <= LML_TOL * np.abs(vref) + 1e-13 * ktt)


def test_c3_full_size_matches_oracle_golden(sm):
    """BASELINE config 3 at FULL size (N = 60 000, D = 784): the fused LML against the oracle value recorded once by
    tests/golden/make_c3_golden.py on the GPU box's host cores (tests/golden/c3_full.json), tolerance 1e-8."""
    import json
    import os
    import torch
    path = os.path.join(os.path.dirname(__file__), "golden", "c3_full.json")
    with open(path) as f:
        g = json.load(f)
    if g.get("oracle_loss") is None:
        pytest.skip("tests/golden/c3_full.json holds no oracle value yet")
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs 30 GB of device memory")
    x, y, *_ = pixel_data(g["n"], g["d"], seed=g["seed"])
    hp, hpd = _hp(sm, **g["hp"])
    sm.device.release_workspaces()
    torch.cuda.empty_cache()
    out, info = sm.device.lml(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), spec=sm.StackSpec(3, "relu", "mlp"),
                              hp=hpd)
    got = float(out[1].item())
    sm.device.release_workspaces()
    torch.cuda.empty_cache()
    assert int(info.item()) == 0
    assert abs(got - g["oracle_loss"]) <= LML_TOL * abs(g["oracle_loss"]), (got, g["oracle_loss"])
    assert abs(float(out[2].item()) - (g["oracle_sum_log_diag"] - 0.5 * g["n"] * np.log(g["hp"]["beta"] / g["hp"]["alpha"]))) \
        <= 1e-9 * abs(g["oracle_sum_log_diag"]) + 1e-6


def test_second_device_in_one_process(sm):
    """single-process multi-device callers (JAX): a call on a tensor of a device that is NOT current must launch there
    (device resolved from the stream / tensor, per-device context: smem attributes, side stream, arena)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    x, y, *_ = regression_data(2500, 8)
    hp, _ = _hp(sm)
    ref = orc.spr_loss(x, y, eps=hp["eps"], kind="student_t", a=hp["alpha"], b=hp["beta"], **_kw(hp, 3, "relu", "mlp"))
    assert torch.cuda.current_device() == 0
    for dev in ("cuda:1", "cuda:0", "cuda:1"):
        hpd = sm.make_hp(device=dev, **hp)
        out, info = sm.device.lml(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev),
                                  spec=sm.StackSpec(3, "relu", "mlp"), hp=hpd)
        assert out.device == torch.device(dev) and int(info.item()) == 0
        assert abs(out[1].item() - ref) <= LML_TOL * abs(ref)
    assert torch.cuda.current_device() == 0
