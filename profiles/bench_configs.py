"""Times every BASELINE.json config that fits one GPU (C1, C2, C4, C5 sweep subset) and checks parity where a
checker finishes in reasonable time (oracle for C1/C2, cuSOLVER-based torch solve for C4).  One JSON line each."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smnngp_b200 as sm
from oracle import nngp_oracle as orc
from tests.synth import regression_data, pixel_data, DEFAULT_HP as HP

def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

def emit(**kw):
    print(json.dumps(kw), flush=True)

hpd = sm.make_hp(**HP)
kw = dict(num_hiddens=3, act="relu", arch="mlp", w_std=HP["w_std"], b_std=HP["b_std"], last_w_std=HP["last_w_std"])
spec = sm.StackSpec(3, "relu", "mlp")
which = sys.argv[1:] or ["c1", "c2", "c4", "c5"]

if "c1" in which or "c2" in which:
    for name, (n, t, d) in (("C1", (404, 52, 13)), ("C1-full", (506, 0, 13)), ("C2", (10000, 1000, 8))):
        if name.startswith("C1") and "c1" not in which: continue
        if name == "C2" and "c2" not in which: continue
        x, y, xt, yt, ym, ys = regression_data(n, d, t=t)
        xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        r = {}
        t_lml = timed(lambda: r.__setitem__("l", sm.device.lml(xd, yd, spec=spec, hp=hpd)))
        ref = orc.spr_loss(x, y, eps=HP["eps"], kind="student_t", a=HP["alpha"], b=HP["beta"], fast=True, **kw)
        out = dict(config=name, N=n, D=d, T=t, lml_ms=t_lml, loss=r["l"][0][1].item(), loss_rel_err=abs(r["l"][0][1].item() - ref) / abs(ref),
                   lml_tflops=(n * (n + 1) * d + n ** 3 / 3 + n * n) / t_lml * 1e-9)
        if t:
            xtd, ytd = torch.from_numpy(xt).cuda(), torch.from_numpy(yt).cuda()
            t_nll = timed(lambda: r.__setitem__("n", sm.device.test_nll(xd, yd, xtd, ytd, ym, ys, spec=spec, hp=hpd)))
            t0 = time.perf_counter()
            ref_nll = orc.spr_test_nll(x, y, xt, yt, ym, ys, eps=HP["eps"], kind="student_t", a=HP["alpha"], b=HP["beta"], **kw)
            out.update(test_nll_ms=t_nll, test_nll=r["n"][0].item(), test_nll_rel_err=abs(r["n"][0].item() - ref_nll) / abs(ref_nll),
                       oracle_test_nll_s=time.perf_counter() - t0)
        emit(**out)

if "c4" in which:
    n, d, t, c = 20000, 3072, 10000, 10
    rng = np.random.default_rng(10)
    x = rng.standard_normal((n, d)); xt = rng.standard_normal((t, d))
    lab = rng.integers(0, c, n); Y = np.eye(c)[lab] - 1.0 / c
    xd, Yd, xtd = torch.from_numpy(x).cuda(), torch.from_numpy(Y).cuda(), torch.from_numpy(xt).cuda()
    r = {}
    t_pred = timed(lambda: r.__setitem__("p", sm.device.predict(xd, Yd, xtd, spec=spec, hp=hpd)), reps=2)
    mean, var, info = r["p"]
    # independent second opinion: cuSOLVER through torch on the (already parity-tested) Gram blocks
    K = sm.device.gram(xd, spec=spec, hp=hpd)
    reg = HP["eps"] * torch.diagonal(K).mean()
    K.diagonal().add_(reg)
    Ktd = sm.device.gram(xtd, xd, spec=spec, hp=hpd)
    L = torch.linalg.cholesky(K)
    mean_ref = Ktd @ torch.cholesky_solve(Yd, L)
    V = torch.linalg.solve_triangular(L, Ktd.T, upper=False)
    ktt = sm.device.nngp_diag(xtd, spec=spec, hp=hpd)
    var_ref = ktt - (V * V).sum(0)
    f_pred = n * (n + 1) * d + 2.0 * t * n * d + n ** 3 / 3 + t * n * float(n) + 2.0 * t * n * (c + 1)
    emit(config="C4", N=n, D=d, T=t, C=c, predict_ms=t_pred, predict_tflops=f_pred / t_pred * 1e-9, info=int(info.item()),
         mean_rel_err_vs_cusolver=float((mean - mean_ref).abs().max() / mean_ref.abs().max()),
         var_rel_err_vs_cusolver=float(((var - var_ref).abs() / var_ref.abs()).max()))
    del K, Ktd, L, V

if "c5" in which:
    for n in (5000, 10000, 20000, 40000):
        x, y, *_ = pixel_data(n, 784)
        xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        for act, b_std in (("relu", 1e-8), ("erf", 0.3)):
            for L in (1, 3, 10):
                hp = sm.make_hp(1.0, b_std, 1.0, 1e-6 if act == "relu" else 1e-4, 2.0, 2.0)
                sp = sm.StackSpec(L, act, "mlp")
                r = {}
                tm = timed(lambda: r.__setitem__("l", sm.device.lml(xd, yd, spec=sp, hp=hp)), reps=2)
                emit(config="C5", N=n, D=784, L=L, act=act, lml_ms=tm, loss=r["l"][0][1].item(), info=int(r["l"][1].item()),
                     tflops=(n * (n + 1) * 784 + n ** 3 / 3 + n * n) / tm * 1e-9)
        del xd, yd
        sm.device.release_workspaces(); torch.cuda.empty_cache()

if "c4draw" in which:
    # config 4 end to end: predictive (shared factorisation) -> S draws per (class, test point) -> NLL / accuracy
    n, d, t, c = 20000, 3072, 10000, 10
    rng = np.random.default_rng(10)
    x = rng.standard_normal((n, d)); xt = rng.standard_normal((t, d))
    lab = rng.integers(0, c, n); Y = np.eye(c)[lab] - 1.0 / c
    labt = rng.integers(0, c, t)
    xd, Yd, xtd = torch.from_numpy(x).cuda(), torch.from_numpy(Y).cuda(), torch.from_numpy(xt).cuda()
    mean, var, info = sm.device.predict(xd, Yd, xtd, spec=spec, hp=hpd)
    for S in (100, 1000, 10000):
        r = {}
        tm = timed(lambda: r.__setitem__("m", sm.device.draw_metrics(mean, var, labt, hp=hpd, num_samples=S, seed=10)), reps=2)
        emit(config="C4-draws", T=t, C=c, S=S, draw_metrics_ms=tm, gdraws_per_s=t * c * S / tm * 1e-6,
             nll=r["m"][0].item(), correct=int(r["m"][1].item()))
