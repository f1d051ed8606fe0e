"""Gradient of SPR.loss w.r.t. the six scalars through the C-ABI against the CPU oracle
and the 50-digit mpmath goldens.

Tolerance: the reference states none for gradients (BASELINE.json: 1e-8 on the LML).  The gradient contracts the
explicit inverse of K + eps I (condition number ~1e6-1e8 at the reference's default eps = 1e-6), so both the
oracle's and the device's values carry ~cond * 2^-53 relative error.  Every component is held to GRAD_TOL relative
to ITSELF (the eps component is 3-5 orders of magnitude larger than the others, so a max-norm would hide them) plus
an absolute floor of 1e-12 of the largest component; the well-conditioned golden cases (eps = 1e-3) to 1e-9."""
import os

import numpy as np
import pytest

from oracle import nngp_oracle as orc
from tests.synth import regression_data, pixel_data, DEFAULT_HP

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRAD_TOL = 1e-7
NAMES = ("w_std", "b_std", "last_w_std", "eps", "alpha", "beta")


@pytest.fixture(scope="module")
def sm():
    import torch
    import smnngp_b200 as s
    assert torch.cuda.is_available()
    s._lib.load()
    return s


def _assert_grad_close(grad, ref, what=""):
    tol = GRAD_TOL * np.abs(ref) + 1e-12 * np.abs(ref).max()
    assert np.all(np.abs(grad - ref) <= tol), f"{what} grad {grad} vs {ref}: err {np.abs(grad - ref)} tol {tol}"


def _hp_vec(hp):
    return np.array([hp[k] for k in NAMES], dtype=np.float64)


def _run(sm, x, y, hp, L, act, arch, kind):
    import torch
    out, grad, info = sm.device.lml_grad(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                                         spec=sm.StackSpec(L, act, arch), hp=torch.from_numpy(_hp_vec(hp)).cuda(),
                                         kind=kind)
    return out.cpu().numpy(), grad.cpu().numpy(), int(info.item())


def _oracle(x, y, hp, L, act, arch, kind):
    return orc.spr_loss_grad(x, y, num_hiddens=L, act=act, arch=arch, w_std=hp["w_std"], b_std=hp["b_std"],
                             last_w_std=hp["last_w_std"], eps=hp["eps"], kind=kind, a=hp["alpha"], b=hp["beta"])


def test_grad_matches_mpmath_golden(sm):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "nngp_golden.npz"))
    x, y = gold["x"], gold["y"]
    for vi, s in enumerate(gold["variants"]):
        act, arch, L, w, b, v = str(s).split(",")
        hp = dict(w_std=float(w), b_std=float(b), last_w_std=float(v), eps=float(gold["eps"]),
                  alpha=float(gold["a"]), beta=float(gold["b"]))
        for kind, key in (("student_t", "t"), ("gauss", "g")):
            out, grad, info = _run(sm, x, y, hp, int(L), act, arch, kind)
            assert info == 0
            assert abs(out[1] - float(gold[f"loss_{key}{vi}"])) <= 1e-10 * abs(out[1])
            ref = gold[f"grad_{key}{vi}"]
            assert np.all(np.abs(grad - ref) <= 1e-9 * np.abs(ref) + 1e-13 * np.abs(ref).max()), (s, kind, grad, ref)


@pytest.mark.parametrize("n,d,L,act,arch,kind,over", [
    (506, 13, 3, "relu", "mlp", "student_t", {}),                      # BASELINE config 1; odd D -> cp.async path
    (506, 13, 3, "relu", "mlp", "gauss", {}),
    (404, 8, 4, "erf", "mlp", "student_t", dict(b_std=0.3)),           # TMA path
    (1300, 8, 2, "relu", "resnet", "student_t", dict(b_std=0.2, w_std=1.3, last_w_std=0.8)),
    (700, 16, 1, "erf", "resnet", "gauss", dict(b_std=0.4)),
    (2600, 8, 3, "relu", "mlp", "student_t", dict(alpha=3.0, beta=1.5)),   # outer panels of 256 + look-ahead
    (77, 5, 2, "relu", "mlp", "gauss", dict(eps=1e-3)),
    (1, 4, 2, "relu", "mlp", "student_t", dict(eps=1e-2)),
    (129, 6, 0, "relu", "mlp", "student_t", dict(eps=1e-2)),           # no hidden layer: linear kernel
])
def test_grad_parity(sm, n, d, L, act, arch, kind, over):
    x, y, *_ = regression_data(n, d)
    hp = dict(DEFAULT_HP)
    hp.update(over)
    ref_loss, ref = _oracle(x, y, hp, L, act, arch, kind)
    out, grad, info = _run(sm, x, y, hp, L, act, arch, kind)
    assert info == 0
    assert abs(out[1] - ref_loss) <= 1e-8 * abs(ref_loss)
    _assert_grad_close(grad, ref)
    # value identical to the value-only entry point (same kernels, same order)
    import torch
    out0, _ = sm.device.lml(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), spec=sm.StackSpec(L, act, arch),
                            hp=torch.from_numpy(_hp_vec(hp)).cuda(), kind=kind)
    assert np.array_equal(out0.cpu().numpy(), out)
    # host-buffer entry point and a second run: bitwise reproducible
    out_h, grad_h, info_h = sm.device.lml_grad(x, y, spec=sm.StackSpec(L, act, arch), hp=_hp_vec(hp), kind=kind)
    assert info_h == 0 and np.array_equal(grad_h, grad) and np.array_equal(out_h, out)


def test_grad_outer_panel_512_pixel_shape(sm):
    """MNIST-shaped inputs, D = 784, forced 512-wide outer panels (the C3 code path) on 3000 points."""
    x, y, *_ = pixel_data(3000, 784)
    hp = dict(DEFAULT_HP)
    hp.update(eps=1e-4)
    sm.device.set_panel_width(512)
    try:
        out, grad, info = _run(sm, x, y, hp, 3, "relu", "mlp", "student_t")
    finally:
        sm.device.set_panel_width(0)
    ref_loss, ref = _oracle(x, y, hp, 3, "relu", "mlp", "student_t")
    assert info == 0
    assert abs(out[1] - ref_loss) <= 1e-8 * abs(ref_loss)
    _assert_grad_close(grad, ref)


def test_grad_non_pd_gives_nan(sm):
    x, y, *_ = regression_data(200, 4)
    x[10] = x[20]
    x[30] = x[20]
    hp = dict(DEFAULT_HP)
    hp.update(eps=1e-300)
    out, grad, info = _run(sm, x, y, hp, 3, "relu", "mlp", "student_t")
    if info != 0:
        assert np.isnan(out[0]) and np.isnan(grad).all()


def test_spr_loss_and_grad_chain_rule(sm):
    """spax mirror: gradients w.r.t. the UNCONSTRAINED variables (what objax.GradValues hands the optimiser,
    regression/train.py:62-66) = d loss / d safe * sigmoid(raw)."""
    x, y, *_ = regression_data(300, 8)
    sp = sm.spax
    kern = sp.NNGPKernel(lambda w, b, v: sm.get_mlp_kernel(3, 1, "relu", w, b, v), w_std=1.2, b_std=0.3, last_w_std=0.9)
    lik = sp.StudentTLikelihood(2.5, 1.5)
    model = sp.SPR(kern, lik, x, y, 0.0, 1.0, eps=1e-4)
    loss, grads = model.loss_and_grad()
    assert abs(loss - model.loss()) <= 1e-14 * abs(loss)
    hp = dict(w_std=1.2, b_std=0.3, last_w_std=0.9, eps=1e-4, alpha=2.5, beta=1.5)
    _, ref = _oracle(x, y, hp, 3, "relu", "mlp", "student_t")
    raw = {k: float(orc.softplus_inverse(hp[k])) for k in NAMES}
    sig = {k: 1.0 / (1.0 + np.exp(-raw[k])) for k in NAMES}
    want = {"kernel.w_std": ref[0] * sig["w_std"], "kernel.b_std": ref[1] * sig["b_std"],
            "kernel.last_w_std": ref[2] * sig["last_w_std"], "eps": ref[3] * sig["eps"],
            "likelihood.a": ref[4] * sig["alpha"], "likelihood.b": ref[5] * sig["beta"]}
    assert set(grads) == set(want)
    scale = max(abs(v) for v in want.values())
    for k in want:
        assert abs(grads[k] - want[k]) <= GRAD_TOL * abs(want[k]) + 1e-12 * scale, (k, grads[k], want[k])


def test_grad_matches_finite_differences_of_the_device_value_at_scale(sm):
    """Size-independent property for sizes the CPU oracle cannot reach (its dual Gram needs ~15 N^2 doubles): the
    analytic gradient must agree with central differences of the DEVICE value (smnngp_lml_f64) in every one of the
    six scalars.  N = 9000 runs the 512-wide outer panels with look-ahead, i.e. the C3 code path."""
    import torch
    n, d = 9000, 64
    x, y, *_ = regression_data(n, d)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    spec = sm.StackSpec(3, "relu", "mlp")
    hp = dict(w_std=1.3, b_std=0.4, last_w_std=0.8, eps=1e-2, alpha=2.5, beta=1.5)

    def value(h):
        out, info = sm.device.lml(xd, yd, spec=spec, hp=torch.from_numpy(_hp_vec(h)).cuda(), kind="student_t")
        assert int(info.item()) == 0
        return float(out[1].item())

    out, grad, info = _run(sm, x, y, hp, 3, "relu", "mlp", "student_t")
    assert info == 0 and out[1] == value(hp)
    for i, name in enumerate(NAMES):
        h = 1e-4 * hp[name]
        up, dn = dict(hp), dict(hp)
        up[name] += h
        dn[name] -= h
        fd = (value(up) - value(dn)) / (2 * h)
        # truncation ~h^2 f''' and round-off ~1e-15 |loss| / h: 1e-5 relative + an absolute floor
        assert abs(grad[i] - fd) <= 2e-5 * abs(fd) + 1e-9, f"{name}: analytic {grad[i]} vs central difference {fd}"
