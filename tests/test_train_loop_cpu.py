"""The training loop of examples/regression_train.py (the reference's experiments/regression/train.py:61-67 loop) and
the spax Adam on CPU: the device calls (loss_and_grad / test_nll) are replaced by the oracle - test infrastructure
only - to check the host-side pieces: variable naming, softplus chain rule, the optimiser, the NaN stop."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import nngp_oracle as orc  # noqa: E402


def _oracle_backed(model, kw_stack):
    """patch the two device entry points of an SPR instance with oracle evaluations (same contracts)"""
    names = ("w_std", "b_std", "last_w_std", "eps", "alpha", "beta")

    def hp():
        ws, bs, ls = model.kernel.get_params()
        d = dict(w_std=ws, b_std=bs, last_w_std=ls, eps=model.eps.safe_value, a=2.0, b=2.0)
        if hasattr(model.likelihood, "a"):
            d.update(a=model.likelihood.a.safe_value, b=model.likelihood.b.safe_value)
        return d

    def loss_and_grad():
        loss, g = orc.spr_loss_grad(model.x_data, model.y_data, kind=model.likelihood.kind, **kw_stack, **hp())
        slots = {"kernel.w_std": (model.kernel.w_std, 0), "kernel.b_std": (model.kernel.b_std, 1),
                 "kernel.last_w_std": (model.kernel.last_w_std, 2), "eps": (model.eps, 3)}
        if hasattr(model.likelihood, "a"):
            slots.update({"likelihood.a": (model.likelihood.a, 4), "likelihood.b": (model.likelihood.b, 5)})
        return loss, {k: float(g[i]) * float(v.constraint.grad(v.value)) for k, (v, i) in slots.items()}

    def test_nll(x, y):
        return orc.spr_test_nll(model.x_data, model.y_data, x, y, model.y_mean, model.y_std, kind=model.likelihood.kind,
                                **kw_stack, **hp())

    model.loss_and_grad, model.test_nll = loss_and_grad, test_nll
    return names


def test_training_loop_decreases_the_loss_and_names_match():
    import importlib.util
    spec = importlib.util.spec_from_file_location("regression_train", os.path.join(ROOT, "examples", "regression_train.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    args = ex.parse(["--method", "tp", "--rows", "120", "--features", "5", "--max-steps", "30", "--print-interval", "10",
                     "--valid-interval", "10", "--epsilon", "1e-3", "--b-std", "0.2", "-lr", "0.05"])
    (x, y), valid, test, (y_std, y_mean) = ex.make_data(args.rows, 15, 15, args.features, args.seed)
    model = ex.build_model(args, x, y, y_mean, y_std)
    assert set(model.vars()) == {"kernel.w_std", "kernel.b_std", "kernel.last_w_std", "eps", "likelihood.a",
                                 "likelihood.b"}
    _oracle_backed(model, dict(num_hiddens=3, act="relu", arch="mlp"))
    l0, g0 = model.loss_and_grad()
    assert set(g0) == set(model.vars())
    lines = []
    best = ex.train(model, args, valid, test, log=lines.append)
    l1, _ = model.loss_and_grad()
    assert l1 < l0 - 1e-3, (l0, l1)                    # 30 Adam steps on the six scalars reduce the training loss
    assert best[0] % 10 == 0 and np.isfinite(best[1])
    assert any("nll:" in s for s in lines) and any("TEST:" in s for s in lines)


def _objax_adam_step(p, m, v, g, lr, step, b1=0.9, b2=0.999, eps=1e-8):
    """objax.optimizer.Adam.__call__ written out by hand: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t),
    m = b1 m + (1 - b1) g, v = b2 v + (1 - b2) g^2, p -= lr_t * m * rsqrt(v + eps)   (eps INSIDE the square root)."""
    import math
    lr_t = lr * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    return p - lr_t * m / math.sqrt(v + eps), m, v


def test_adam_matches_objax_update_rule():
    import math
    from smnngp_b200.spax import Adam, ConstraintTrainVar, positive
    # (a) ordinary gradient, (b) the tiny gradient of b_std at the reference defaults (|g| ~ 1e-8): with eps inside the
    # square root the step is ~lr * 1e-4, with sqrt(v) + eps it would be ~lr - orders of magnitude apart
    for grads in ([2.0, -1.5, 0.7], [5e-8, 3e-8, -4e-8]):
        v = ConstraintTrainVar(1.5, constraint=positive())
        opt = Adam({"v": v})
        p, m, vv = float(v.value), 0.0, 0.0
        for step, g in enumerate(grads, 1):
            opt(0.1, {"v": g})
            p, m, vv = _objax_adam_step(p, m, vv, g, 0.1, step)
            assert abs(float(v.value) - p) <= 1e-15 * max(1.0, abs(p)), (step, float(v.value), p)
    v = ConstraintTrainVar(1.5, constraint=positive())
    raw0 = float(v.value)
    Adam({"v": v})(0.1, {"v": 5e-8})
    assert abs(float(v.value) - raw0) < 0.1 * 1e-3            # effectively frozen, as in the reference
    # NaN gradients propagate like in objax; skipping them is an explicit opt-in
    v = ConstraintTrainVar(1.5, constraint=positive())
    Adam({"v": v})(0.1, {"v": float("nan")})
    assert math.isnan(float(v.value))
    v = ConstraintTrainVar(1.5, constraint=positive())
    opt = Adam({"v": v}, skip_nan=True)
    opt(0.1, {"v": float("nan")})
    assert float(v.value) == raw0 and opt.step == 0
    g = float(positive().grad(v.value))
    h = 1e-6
    assert abs(g - (float(positive()(v.value + h)) - float(positive()(v.value - h))) / (2 * h)) <= 1e-8


def test_distributed_spr_glue_with_injected_solvers():
    """DistributedSPR host logic without GPUs: the solvers are injected (oracle-backed stand-ins with the interface of
    distributed.DistributedGrad / DistributedPredict); the softplus chain rule and the variable naming must match SPR's"""
    import torch
    import smnngp_b200 as sm
    from oracle import nngp_oracle as orc
    from smnngp_b200.spax import NNGPKernel, StudentTLikelihood, DistributedSPR
    from tests.synth import regression_data, DEFAULT_HP as hp
    x, y, xt, yt, ym, ys = regression_data(300, 5, t=40)
    kw = dict(num_hiddens=2, act="relu", arch="mlp")

    class GradStub:
        def lml_grad(self, xd, yd, hpd, kind="student_t"):
            h = hpd.tolist()
            loss, g = orc.spr_loss_grad(xd.numpy(), yd.numpy(), w_std=h[0], b_std=h[1], last_w_std=h[2], eps=h[3], kind=kind,
                                        a=h[4], b=h[5], **kw)
            return torch.tensor([0.0, loss, 0.0, 0.0], dtype=torch.float64), torch.from_numpy(np.asarray(g)), torch.zeros(1)

    class PredictStub:
        def test_nll(self, xd, yd, xtd, ytd, y_mean, y_std, hpd, kind="student_t"):
            h = hpd.tolist()
            nll = orc.spr_test_nll(xd.numpy(), yd.numpy(), xtd.numpy(), ytd.numpy(), y_mean, y_std, w_std=h[0], b_std=h[1],
                                   last_w_std=h[2], eps=h[3], kind=kind, a=h[4], b=h[5], **kw)
            return torch.tensor([nll], dtype=torch.float64), None, None, torch.zeros(1)

    def get_kernel_fn(w_std, b_std, last_w_std):
        return sm.get_mlp_kernel(2, act="relu", w_std=w_std, b_std=b_std, last_w_std=last_w_std)

    model = DistributedSPR(NNGPKernel(get_kernel_fn, 1.2, 0.3, 0.8), StudentTLikelihood(2.5, 1.5), torch.from_numpy(x),
                           torch.from_numpy(y), ym, ys, eps=1e-3, solvers={"grad": GradStub(), "predict": {40: PredictStub()}})
    loss, grads = model.loss_and_grad()
    assert set(grads) == set(model.vars())
    lref, gref = orc.spr_loss_grad(x, y, w_std=1.2, b_std=0.3, last_w_std=0.8, eps=1e-3, kind="student_t", a=2.5, b=1.5, **kw)
    assert abs(loss - lref) <= 1e-12 * abs(lref)
    # chain rule through softplus: d loss / d raw = d loss / d safe * sigmoid(raw)
    raw = float(model.kernel.w_std.value)
    assert abs(grads["kernel.w_std"] - gref[0] / (1.0 + np.exp(-raw))) <= 1e-10 * abs(gref[0])
    nll = float(model.test_nll(torch.from_numpy(xt), torch.from_numpy(yt)))
    ref = orc.spr_test_nll(x, y, xt, yt, ym, ys, w_std=1.2, b_std=0.3, last_w_std=0.8, eps=1e-3, kind="student_t", a=2.5,
                           b=1.5, **kw)
    assert abs(nll - ref) <= 1e-12 * abs(ref)
